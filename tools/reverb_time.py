"""Times effects.Reverb-shaped convolutions (48000-tap IR on 64000 samples): the
hand-written partitioned overlap-save kernels vs the framed cuFFT formulation."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import core
rng = np.random.default_rng(0)
for B, shared in ((32, False), (32, True), (256, True)):
  audio = torch.from_numpy(rng.standard_normal((B, 64000)).astype(np.float32)).cuda()
  ir = torch.from_numpy((rng.standard_normal((1 if shared else B, 48000)) *
                         np.exp(-np.arange(48000) / 8000.0)).astype(np.float32)).cuda()
  def ours():
    return core.fft_convolve(audio, ir, padding='same', delay_compensation=0)
  def cufft():
    fft_size = core.get_fft_size(64000, 48000)
    return core._fft_convolve_cufft(audio, ir[:, None, :], 1, 64000, fft_size, 0, 64000)
  for name, fn in (('partitioned overlap-save (ours)', ours), ('framed cuFFT', cufft)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print('B=%d shared_ir=%s %-34s %.3f ms' % (B, shared, name, e0.elapsed_time(e1) / 10), flush=True)
  d = (ours() - cufft()).abs().max().item()
  print('  max |ours - cufft| = %.2e' % d)
