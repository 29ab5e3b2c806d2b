"""Times synths.Sinusoidal.get_signal: the fused frame-rate oscillator bank against the
reference's decomposition (resample + resample + oscillator_bank over [B, N, K])."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import core
rng = np.random.default_rng(0)
for B, K in ((8, 100), (32, 100)):
  F, N = 1000, 64000
  freqs = torch.from_numpy(rng.uniform(50, 7000, (B, F, K)).astype(np.float32)).cuda()
  amps = torch.from_numpy(rng.uniform(0, 1, (B, F, K)).astype(np.float32)).cuda()
  def fused():
    return core.sinusoidal_synthesis(freqs, amps, n_samples=N)
  def materialised():
    return core.oscillator_bank(core.resample(freqs, N), core.resample(amps, N, method='window'))
  for name, fn in (('fused frame-rate bank', fused), ('resample + resample + oscillator_bank', materialised)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print('B=%d K=%d %-40s %.3f ms' % (B, K, name, e0.elapsed_time(e1) / 5), flush=True)
  print('  max |fused - materialised| = %.2e' % (fused() - materialised()).abs().max().item())
