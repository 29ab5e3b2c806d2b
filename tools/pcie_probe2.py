import os, sys, time, torch
dev = torch.device('cuda')
n1, n2 = int(21.4e6 / 4), int(8.2e6 / 4)
h1 = torch.empty(n1).pin_memory(); h1.fill_(1.0); d1 = torch.empty(n1, device=dev)
h2 = torch.empty(n2).pin_memory(); d2 = torch.ones(n2, device=dev)
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def warm(t=0.3):
  t0 = time.perf_counter()
  while time.perf_counter() - t0 < t:
    d1.copy_(h1, non_blocking=True); h2.copy_(d2, non_blocking=True)
  torch.cuda.synchronize()
def run(name, fn, n=60):
  warm()
  t0 = time.perf_counter()
  for _ in range(n): fn()
  torch.cuda.synchronize()
  print('%-50s %.1f us' % (name, (time.perf_counter() - t0) / n * 1e6), flush=True)
def h2d_nosync(): d1.copy_(h1, non_blocking=True)
def h2d_sync(): d1.copy_(h1, non_blocking=True); torch.cuda.current_stream().synchronize()
def d2h_sync(): h2.copy_(d2, non_blocking=True); torch.cuda.current_stream().synchronize()
def both_sync():
  with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
  with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
  s1.synchronize(); s2.synchronize()
def h2d_split_sync():
  k = n1 * 6 // 10
  with torch.cuda.stream(s1): d1[:k].copy_(h1[:k], non_blocking=True)
  with torch.cuda.stream(s3): d1[k:].copy_(h1[k:], non_blocking=True)
  s1.synchronize(); s3.synchronize()
def h2d_8_sync():
  k = n1 // 8
  for i in range(8): d1[i*k:(i+1)*k].copy_(h1[i*k:(i+1)*k], non_blocking=True)
  torch.cuda.current_stream().synchronize()
def h2d_then_d2h_sync():
  d1.copy_(h1, non_blocking=True); h2.copy_(d2, non_blocking=True); torch.cuda.current_stream().synchronize()
run('H2D 21.4MB queued back to back', h2d_nosync)
run('H2D 21.4MB + sync', h2d_sync)
run('D2H 8.2MB + sync', d2h_sync)
run('H2D || D2H (2 streams) + sync', both_sync)
run('H2D split over 2 streams + sync', h2d_split_sync)
run('H2D as 8 copies on one stream + sync', h2d_8_sync)
run('H2D then D2H one stream + sync', h2d_then_d2h_sync)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_b200
from ddsp_b200 import host
from tests.util import synth_inputs
B = 32
inp = synth_inputs(B, 1000, 100, 65, 64000, seed=1234)
keys = ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']
pinned = {k: host.pin(inp[k]) for k in keys}
out = host.pinned_empty((B, 64000))
group = ddsp_b200.ProcessorGroup(dag=[
    (ddsp_b200.Harmonic(), ['amps', 'harmonic_distribution', 'f0_hz']),
    (ddsp_b200.FilteredNoise(window_size=0), ['noise_magnitudes']),
    (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
for chunks in (1, 2, 3, 4, 5, 6):
  dec = ddsp_b200.HostDecoder(group, B, 1000, 100, 65, n_chunks=chunks)
  run('HostDecoder chunks=%d (probe warm-up)' % chunks, lambda: dec(pinned, out=out))
  def selfwarm():
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.3: dec(pinned, out=out)
    t0 = time.perf_counter()
    for _ in range(60): dec(pinned, out=out)
    print('%-50s %.1f us' % ('HostDecoder chunks=%d (self warm-up 0.3 s)' % chunks, (time.perf_counter() - t0) / 60 * 1e6), flush=True)
  selfwarm()
  dec.close()
