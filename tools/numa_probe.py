import os, glob, subprocess, torch
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
try:
  print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout)
except Exception as e: print(e)
for n in sorted(glob.glob('/sys/devices/system/node/node*')):
  try: print(n, open(n + '/cpulist').read().strip())
  except Exception as e: print(n, e)
p = torch.cuda.get_device_properties(0)
bus = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
print('gpu0 pci', bus)
try: print('numa_node', open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
except Exception as e: print('numa_node ?', e)
print('my cpu', os.sched_getcpu() if hasattr(os, 'sched_getcpu') else '?')
