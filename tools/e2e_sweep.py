"""Times HostDecoder (host buffers in/out) for several chunk counts."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import host
from tests.util import synth_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
keys = ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']
pinned = {k: host.pin(inp[k]) for k in keys}
out = host.pinned_empty((B, 64000))
group = ddsp_b200.ProcessorGroup(dag=[
    (ddsp_b200.Harmonic(), ['amps', 'harmonic_distribution', 'f0_hz']),
    (ddsp_b200.FilteredNoise(window_size=0), ['noise_magnitudes']),
    (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
for chunks in (1, 2, 3, 4, 5, 6, 8):
  dec = ddsp_b200.HostDecoder(group, B, 1000, 100, 65, n_chunks=chunks)
  for _ in range(150): dec(pinned, out=out)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(30): dec(pinned, out=out)
  e1.record(); torch.cuda.synchronize()
  print('B=%d chunks=%d e2e %.1f us/step' % (B, chunks, 1e3 * e0.elapsed_time(e1) / 30), flush=True)
  dec.close()
