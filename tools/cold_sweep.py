"""Harmonic kernel / decoder step with HBM-cold inputs (ring of input sets > 2x L2),
for the current DDSP_B200_HARM_FW knob."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import _lib, core
from tests.util import synth_inputs
lib = _lib.load()
tag = 'FW=%s' % os.environ.get('DDSP_B200_HARM_FW', 'auto')
for B in (32, 256):
  inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
  keys = ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']
  set_bytes = sum(inp[k].nbytes for k in keys) + 4 * B * 64000
  n_sets = max(2, -(-2 * 126 * 2**20 // set_bytes))
  sets = [{k: torch.from_numpy(inp[k]).cuda() + 0.001 * s for k in keys} for s in range(n_sets)]
  for s in sets: s['f0_hz'] = torch.from_numpy(inp['f0_hz']).cuda()
  ctl = [ddsp_b200.Harmonic().get_controls(s['amps'], s['harmonic_distribution'], s['f0_hz']) for s in sets]
  outs = [torch.empty(B, 64000, device='cuda') for _ in range(n_sets)]
  st = torch.cuda.current_stream().cuda_stream
  def harm(i):
    c = ctl[i % n_sets]
    _lib.check(lib.ddsp_b200_harmonic_forward(c['f0_hz'].data_ptr(), c['amplitudes'].data_ptr(),
        c['harmonic_distribution'].data_ptr(), outs[i % n_sets].data_ptr(), B, 1000, 100, 64000, 16000.0, 0, 0, 0, st))
  def dec(i):
    s = sets[i % n_sets]
    core.decoder_forward(s['amps'], s['harmonic_distribution'], s['f0_hz'], s['noise_magnitudes'], 64000)
  for name, fn in (('harmonic_forward', harm), ('decoder step', dec)):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60): fn(10 + i)
    e1.record(); torch.cuda.synchronize()
    print('%s cold B=%d %s %.1f us' % (tag, B, name, 1e3 * e0.elapsed_time(e1) / 60), flush=True)
