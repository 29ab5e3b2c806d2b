"""PCIe probe: H2D / D2H bandwidth from pinned memory, alone and concurrently, and
a torch-stream emulation of the chunked host pipeline (for comparison with
ddsp_b200_decoder_forward_host)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import core, host
from tests.util import synth_inputs

dev = torch.device('cuda')
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=20, warm=3):
  for _ in range(warm): fn()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(n): fn()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / n * 1e6

for mb in (1, 4, 16, 21.4):
  n = int(mb * 1e6 / 4)
  h = torch.empty(n).pin_memory(); d = torch.empty(n, device=dev)
  def h2d(): d.copy_(h, non_blocking=True); torch.cuda.current_stream().synchronize()
  def d2h(): h.copy_(d, non_blocking=True); torch.cuda.current_stream().synchronize()
  a, b = timeit(h2d), timeit(d2h)
  print('%.1f MB  H2D %.1f us (%.1f GB/s)   D2H %.1f us (%.1f GB/s)' % (mb, a, mb * 1e3 / a, b, mb * 1e3 / b), flush=True)

# concurrent H2D + D2H on two streams
n1, n2 = int(21.4e6 / 4), int(8.2e6 / 4)
h1 = torch.empty(n1).pin_memory(); d1 = torch.empty(n1, device=dev)
h2 = torch.empty(n2).pin_memory(); d2 = torch.empty(n2, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
  with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
  with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
  s1.synchronize(); s2.synchronize()
print('concurrent 21.4 MB H2D + 8.2 MB D2H: %.1f us' % timeit(both), flush=True)

# torch-stream emulation of the chunked pipeline
B = 32
inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
keys = ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']
pinned = {k: host.pin(inp[k]) for k in keys}
devbuf = {k: torch.empty_like(pinned[k], device=dev) for k in keys}
out_d = torch.empty(B, 64000, device=dev)
out_h = host.pinned_empty((B, 64000))
sh, sc, sd = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def pipeline(chunks):
  per = (B + chunks - 1) // chunks
  for b0 in range(0, B, per):
    b1 = min(B, b0 + per)
    with torch.cuda.stream(sh):
      for k in keys: devbuf[k][b0:b1].copy_(pinned[k][b0:b1], non_blocking=True)
      e1 = torch.cuda.Event(); e1.record(sh)
    with torch.cuda.stream(sc):
      sc.wait_event(e1)
      a = core.decoder_forward(devbuf['amps'][b0:b1], devbuf['harmonic_distribution'][b0:b1],
                               devbuf['f0_hz'][b0:b1], devbuf['noise_magnitudes'][b0:b1], 64000)
      e2 = torch.cuda.Event(); e2.record(sc)
    with torch.cuda.stream(sd):
      sd.wait_event(e2)
      out_h[b0:b1].copy_(a, non_blocking=True)
  sd.synchronize()
for chunks in (1, 2, 4, 8):
  print('torch-stream pipeline chunks=%d: %.1f us' % (chunks, timeit(lambda: pipeline(chunks))), flush=True)

group = ddsp_b200.ProcessorGroup(dag=[
    (ddsp_b200.Harmonic(), ['amps', 'harmonic_distribution', 'f0_hz']),
    (ddsp_b200.FilteredNoise(window_size=0), ['noise_magnitudes']),
    (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
for chunks in (1, 2, 3, 4, 8):
  dec = ddsp_b200.HostDecoder(group, B, 1000, 100, 65, n_chunks=chunks)
  print('C pipeline chunks=%d: %.1f us' % (chunks, timeit(lambda: dec(pinned, out=out_h))), flush=True)
  dec.close()
