"""H2D / D2H bandwidth of a pinned buffer allocated with the thread bound to each
NUMA node (the GPU sits on one of them)."""
import os, sys, time, glob, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_b200 import host
dev = torch.device('cuda')
torch.zeros(1, device=dev)
def bw(h, d, n=30):
  for _ in range(3): d.copy_(h, non_blocking=True)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): d.copy_(h, non_blocking=True)
  torch.cuda.synchronize(); a = (time.perf_counter() - t0) / n
  for _ in range(3): h.copy_(d, non_blocking=True)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): h.copy_(d, non_blocking=True)
  torch.cuda.synchronize(); b = (time.perf_counter() - t0) / n
  return h.numel() * 4 / a / 1e9, h.numel() * 4 / b / 1e9
all_cpus = os.sched_getaffinity(0)
n = int(21.4e6 / 4)
d = torch.empty(n, device=dev)
for node in sorted(glob.glob('/sys/devices/system/node/node[0-9]*')):
  cpus = host._cpulist(open(node + '/cpulist').read()) & all_cpus
  os.sched_setaffinity(0, cpus)
  h = torch.empty(n).pin_memory(); h.fill_(1.0)
  for rep in range(3):
    print(os.path.basename(node), 'H2D %.1f GB/s  D2H %.1f GB/s' % bw(h, d), flush=True)
  del h
os.sched_setaffinity(0, all_cpus)
print('bind_to_device_numa_node ->', host.bind_to_device_numa_node())
h = torch.empty(n).pin_memory(); h.fill_(1.0)
for rep in range(3):
  print('bound: H2D %.1f GB/s  D2H %.1f GB/s' % bw(h, d), flush=True)
