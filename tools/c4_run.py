"""One configs[3] step (decoder fwd + bwd through SpectralLoss, B=128) for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import autograd as ag, losses
from tests.util import synth_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
inp = synth_inputs(B, 1000, 100, 65, 64000, seed=55)
d = {k: torch.from_numpy(inp[k]).cuda() for k in ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']}
for k in ('amps', 'harmonic_distribution', 'noise_magnitudes'): d[k].requires_grad_(True)
target = 0.1 * torch.randn(B, 64000, device='cuda')
loss_obj = losses.SpectralLoss(mag_weight=1.0, logmag_weight=1.0)
def step(i):
  for k in ('amps', 'harmonic_distribution', 'noise_magnitudes'): d[k].grad = None
  audio = ag.decoder_train(d['amps'], d['harmonic_distribution'], d['f0_hz'], d['noise_magnitudes'],
                           n_samples=64000, window_size=0, seed=1, offset=i)
  loss_obj(target, audio).backward()
for i in range(2): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(2); e1.record(); torch.cuda.synchronize()
print('c4 step %.2f ms' % e0.elapsed_time(e1))
