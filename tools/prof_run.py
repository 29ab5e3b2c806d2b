"""Runs each kernel of the decoder a few times at a given batch (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200  # noqa: E402
from tests.util import synth_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
feats = {k: torch.from_numpy(inp[k]).cuda() for k in
         ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']}
harm = ddsp_b200.Harmonic()
noise = ddsp_b200.FilteredNoise(window_size=0)
group = ddsp_b200.ProcessorGroup(dag=[
    (harm, ['amps', 'harmonic_distribution', 'f0_hz']),
    (noise, ['noise_magnitudes']),
    (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
for _ in range(reps):
  audio = group(feats)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(reps):
  audio = group(feats)
ev[1].record()
torch.cuda.synchronize()
print('B=%d step %.3f ms' % (B, ev[0].elapsed_time(ev[1]) / reps))
