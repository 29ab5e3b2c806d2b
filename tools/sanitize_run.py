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
# round-2 kernels: harmonic_shifts / fused sinusoidal bank, cubic + 4-D resample,
# angular_cumsum (exact and TF order), tf_sequential bank, long-IR convolution forward
# and backward, d f0, f0 < 1 Hz frames
shifts = 0.01 * torch.randn(B, F, K, device='cuda')
h2 = core.harmonic_synthesis(feats['f0_hz'], outs['harmonic']['controls']['amplitudes'],
                             harmonic_shifts=shifts,
                             harmonic_distribution=outs['harmonic']['controls']['harmonic_distribution'],
                             n_samples=N)
r4 = core.resample(torch.randn(2, 9, 3, 2), 37, method='cubic', add_endpoint=False)
om = 0.3 * torch.rand(2, 2300, 3, device='cuda')
pc = core.angular_cumsum(om); ps = core.angular_cumsum(om, tf_sequential=True)
ob = core.oscillator_bank(200 + 3000 * torch.rand(2, 700, 4, device='cuda'),
                          torch.rand(2, 700, 4, device='cuda'), phase_mode='tf_sequential',
                          use_angular_cumsum=True)
xa = torch.randn(2, 5000, device='cuda', requires_grad=True)
hi = (0.02 * torch.randn(1, 4500, device='cuda')).requires_grad_(True)
core.fft_convolve(xa, hi, padding='same', delay_compensation=0).square().mean().backward()
f0g = feats['f0_hz'].clone()
f0g[:, 5:9] = 0.3
f0g.requires_grad_(True)
ag.decoder_train(feats['amps'], feats['harmonic_distribution'], f0g,
                 feats['noise_magnitudes'], n_samples=N, window_size=0).abs().mean().backward()
torch.cuda.synchronize()
assert torch.isfinite(h2).all() and torch.isfinite(r4).all() and torch.isfinite(ob).all()
assert torch.isfinite(xa.grad).all() and torch.isfinite(hi.grad).all() and torch.isfinite(f0g.grad).all()
assert torch.isfinite(e).all() and torch.isfinite(f).all() and torch.isfinite(g).all()
print('sanitize_run ok', float(a.abs().mean()), float(b.abs().mean()),
      float(c.abs().mean()), float(d.abs().mean()),
      float(raw['harmonic_distribution'].grad.abs().mean()))
