"""Times the harmonic kernel alone at B=256 / B=32 in the two ways it is used:
stand-alone on Harmonic.get_controls' output, and as the first kernel of the
decoder (raw network outputs, get_controls fused) - the second by timing the whole
decoder step and the noise kernel alone (accumulate) and taking the difference -
plus the decoder step itself.  Knobs: DDSP_B200_HARM_FW."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import _lib, core
from tests.util import synth_inputs

lib = _lib.load()
tag = 'FW=%s %s' % (os.environ.get('DDSP_B200_HARM_FW', 'auto'), os.environ.get('TAG', ''))
for B in (256, 32):
  sets = []
  for s in range(3):
    inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234 + s)
    f = {k: torch.from_numpy(inp[k]).cuda() for k in ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']}
    ctl = ddsp_b200.Harmonic().get_controls(f['amps'], f['harmonic_distribution'], f['f0_hz'])
    nctl = ddsp_b200.FilteredNoise().get_controls(f['noise_magnitudes'])
    sets.append((f, ctl, nctl, torch.empty(B, 64000, device='cuda')))
  st = torch.cuda.current_stream().cuda_stream

  def harm(i):
    f, ctl, nctl, out = sets[i % 3]
    _lib.check(lib.ddsp_b200_harmonic_forward(
        ctl['f0_hz'].data_ptr(), ctl['amplitudes'].data_ptr(), ctl['harmonic_distribution'].data_ptr(),
        out.data_ptr(), B, 1000, 100, 64000, 16000.0, 0, 0, 0, st))

  def noise(i):
    f, ctl, nctl, out = sets[i % 3]
    _lib.check(lib.ddsp_b200_filtered_noise_forward(
        nctl['magnitudes'].data_ptr(), None, 7, i, out.data_ptr(), B, 1000, 65, 64000, 0, 1, None, 0, st))

  def dec(i):
    f = sets[i % 3][0]
    return core.decoder_forward(f['amps'], f['harmonic_distribution'], f['f0_hz'], f['noise_magnitudes'], 64000)

  res = {}
  for name, fn in (('harmonic_forward', harm), ('noise(accumulate)', noise), ('decoder step', dec)):
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(24): fn(i)
    e1.record(); torch.cuda.synchronize()
    res[name] = 1e3 * e0.elapsed_time(e1) / 24
    print('%s B=%d %s %.1f us' % (tag, B, name, res[name]), flush=True)
