"""Times harmonic_forward alone for the current DDSP_B200_HARM_NT / _FT knobs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import core
from tests.util import synth_inputs
for B in (256, 32):
  inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
  f = {k: torch.from_numpy(inp[k]).cuda() for k in ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']}
  def run():
    return core.decoder_forward(f['amps'], f['harmonic_distribution'], f['f0_hz'], f['noise_magnitudes'], 64000)
  for _ in range(5): run()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): run()
  e1.record(); torch.cuda.synchronize()
  print('NT=%s FT=%s B=%d decoder step %.1f us' % (os.environ.get('DDSP_B200_HARM_NT'), os.environ.get('DDSP_B200_HARM_FT'), B, 1e3 * e0.elapsed_time(e1) / 20))
