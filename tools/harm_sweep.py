"""Times the harmonic kernel alone (raw inputs, get_controls fused) for the
current DDSP_B200_HARM_IMPL / DDSP_B200_HARM_FW knobs, plus the decoder step."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import _lib, core
from tests.util import synth_inputs

lib = _lib.load()
tag = 'harmonic_v4 FW=%s' % os.environ.get('DDSP_B200_HARM_FW', 'auto')
for B in (256, 32):
  inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
  f = {k: torch.from_numpy(inp[k]).cuda() for k in ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']}
  out = torch.empty(B, 64000, device='cuda')
  ctl = ddsp_b200.Harmonic().get_controls(f['amps'], f['harmonic_distribution'], f['f0_hz'])
  st = torch.cuda.current_stream().cuda_stream

  def harm():   # controls -> audio (no fused get_controls)
    _lib.check(lib.ddsp_b200_harmonic_forward(
        ctl['f0_hz'].data_ptr(), ctl['amplitudes'].data_ptr(), ctl['harmonic_distribution'].data_ptr(),
        out.data_ptr(), B, 1000, 100, 64000, 16000.0, 0, 0, 0, st))

  def dec():
    return core.decoder_forward(f['amps'], f['harmonic_distribution'], f['f0_hz'], f['noise_magnitudes'], 64000)

  for name, fn in (('harmonic_forward', harm), ('decoder step', dec)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print('%s B=%d %s %.1f us' % (tag, B, name, 1e3 * e0.elapsed_time(e1) / 20), flush=True)
