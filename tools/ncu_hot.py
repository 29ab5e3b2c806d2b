"""Summarise an `ncu --page source --csv` dump: hot SASS regions by executed
instructions and stall samples.  usage: ncu_hot.py file.csv [chunk] [kernel#]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 50
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
# split into per-kernel blocks (each starts with a "Kernel Name" row)
blocks, cur = [], None
for r in rows:
  if r and r[0] == 'Kernel Name':
    cur = {'name': r[1], 'rows': []}
    blocks.append(cur)
  elif cur is not None:
    cur['rows'].append(r)
blk = blocks[which]
hdr = blk['rows'][0]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk['rows'][1:] if len(r) >= len(hdr) - 5]


def gi(r, k):
  try:
    return int(float(r[idx[k]]))
  except (ValueError, KeyError, IndexError):
    return 0


tot_i = sum(gi(r, 'Instructions Executed') for r in data)
tot_s = sum(gi(r, '# Samples') for r in data)
print(blk['name'][:80], 'kernels in file:', len(blocks))
print('rows', len(data), 'inst', tot_i, 'samples', tot_s)
for s in range(0, len(data), chunk):
  b = data[s:s + chunk]
  ins = sum(gi(r, 'Instructions Executed') for r in b)
  sm = sum(gi(r, '# Samples') for r in b)
  ex = sum(gi(r, 'L1 Wavefronts Shared Excessive') for r in b)
  if ins > 0.01 * tot_i or sm > 0.01 * tot_s:
    ops = {}
    for r in b:
      t = r[idx['Source']].split()
      op = t[0] if t else ''
      if op.startswith('@') and len(t) > 1:
        op = t[1]
      ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print('rows %5d-%5d inst %5.1f%% samples %5.1f%% excess_smem_wavefronts %9d %s' %
          (s, s + chunk, 100.0 * ins / tot_i, 100.0 * sm / max(tot_s, 1), ex, top))
if len(sys.argv) > 5:
  a, bnd = int(sys.argv[4]), int(sys.argv[5])
  for i in range(a, bnd):
    r = data[i]
    print(i, '%10d %5d %7d' % (gi(r, 'Instructions Executed'), gi(r, '# Samples'),
                               gi(r, 'L1 Wavefronts Shared Excessive')),
          r[idx['Source']][:100])
