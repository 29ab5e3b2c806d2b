"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck).

  compute-sanitizer --tool memcheck  python tools/sanitize_run.py
  compute-sanitizer --tool racecheck python tools/sanitize_run.py
  compute-sanitizer --tool synccheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200  # noqa: E402
from ddsp_b200 import autograd as ag, core  # noqa: E402
from tests.util import synth_inputs  # noqa: E402

B, F, K, nb, N = 3, 70, 100, 65, 70 * 64      # 3 tiles per item, edges included
inp = synth_inputs(B, F, K, nb, N, seed=1)
feats = {k: torch.from_numpy(inp[k]).cuda() for k in
         ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']}
harm = ddsp_b200.Harmonic(n_samples=N)
noise = ddsp_b200.FilteredNoise(n_samples=N, window_size=0)
group = ddsp_b200.ProcessorGroup(dag=[
    (harm, ['amps', 'harmonic_distribution', 'f0_hz']),
    (noise, ['noise_magnitudes']),
    (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
a = group(feats)                                 # decoder_forward (fast + pipelined)
outs = group.get_controls(feats)                 # per-processor kernels + Add
b = core.filtered_noise(outs['filtered_noise']['controls']['magnitudes'][:, :, :33]
                        .contiguous(), N, window_size=0)          # generic fused kernel
c = core.frequency_filter(torch.rand(2, 1000).cuda(), torch.rand(2, 13, 513).cuda(),
                          window_size=257)       # generic ir + fir kernels
d, ph = core.streaming_harmonic_synthesis(feats['f0_hz'][:, :2], torch.rand(B, 2, 1).cuda(),
                                          torch.rand(B, 2, K).cuda(), n_samples=320)
raw = {k: feats[k].clone().requires_grad_(True) for k in
       ['amps', 'harmonic_distribution', 'noise_magnitudes']}
audio = ag.decoder_train(raw['amps'], raw['harmonic_distribution'], feats['f0_hz'],
                         raw['noise_magnitudes'], n_samples=N, window_size=0)
audio.square().mean().backward()
# fused spectral-loss pieces, host-buffer pipeline, Sinusoidal, Reverb (cuFFT route)
from ddsp_b200 import losses, host
tgt = 0.1 * torch.randn(B, N, device='cuda')
aud = (0.1 * torch.randn(B, N, device='cuda')).requires_grad_(True)
losses.SpectralLoss(fft_sizes=(256, 64), mag_weight=1.0, logmag_weight=1.0)(tgt, aud).backward()
dec = ddsp_b200.HostDecoder(group, B, F, K, nb, n_chunks=2)
e = dec({k: host.pin(inp[k]) for k in ['amps', 'harmonic_distribution', 'f0_hz',
                                        'noise_magnitudes']})
dec.close()
f = ddsp_b200.Sinusoidal(n_samples=640)(torch.randn(2, 10, 6), torch.randn(2, 10, 6))
g = ddsp_b200.Reverb()(torch.randn(2, 3000).cuda(), 0.01 * torch.randn(2, 2500).cuda())
torch.cuda.synchronize()
assert torch.isfinite(e).all() and torch.isfinite(f).all() and torch.isfinite(g).all()
print('sanitize_run ok', float(a.abs().mean()), float(b.abs().mean()),
      float(c.abs().mean()), float(d.abs().mean()),
      float(raw['harmonic_distribution'].grad.abs().mean()))
