import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import core
from tests.util import synth_inputs
B = 32
inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
f = {k: torch.from_numpy(inp[k]).cuda() for k in ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']}
group = ddsp_b200.ProcessorGroup(dag=[
    (ddsp_b200.Harmonic(), ['amps', 'harmonic_distribution', 'f0_hz']),
    (ddsp_b200.FilteredNoise(window_size=0), ['noise_magnitudes']),
    (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
def ev_time(fn, n=200):
  for _ in range(20): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter(); e0.record()
  for _ in range(n): fn()
  e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
  return 1e3 * e0.elapsed_time(e1) / n, (t1 - t0) / n * 1e6
print('eager group():   gpu %.1f us/step, cpu submit %.1f us/step' % ev_time(lambda: group(f)))
print('eager decoder_forward: gpu %.1f us/step, cpu %.1f' % ev_time(lambda: core.decoder_forward(f['amps'], f['harmonic_distribution'], f['f0_hz'], f['noise_magnitudes'], 64000)))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
  for _ in range(3): out = core.decoder_forward(f['amps'], f['harmonic_distribution'], f['f0_hz'], f['noise_magnitudes'], 64000)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
  out = core.decoder_forward(f['amps'], f['harmonic_distribution'], f['f0_hz'], f['noise_magnitudes'], 64000)
print('graph replay:    gpu %.1f us/step, cpu %.1f' % ev_time(g.replay))
ref = core.decoder_forward(f['amps'], f['harmonic_distribution'], f['f0_hz'], f['noise_magnitudes'], 64000, offset=0)
