"""Per-region stall reasons of one kernel from an `ncu --page source --csv` dump.
usage: ncu_stalls.py file.csv kernel# [chunk]   (regions with > 1 % of the samples)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
blocks, cur = [], None
for r in rows:
  if r and r[0] == 'Kernel Name':
    cur = {'name': r[1], 'rows': []}; blocks.append(cur)
  elif cur is not None:
    cur['rows'].append(r)
blk = blocks[which]; hdr = blk['rows'][0]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk['rows'][1:] if len(r) >= len(hdr) - 5]
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def gi(r, k):
  try: return int(float(r[idx[k]]))
  except (ValueError, KeyError, IndexError): return 0
tot = sum(gi(r, '# Samples') for r in data); toti = sum(gi(r, 'Instructions Executed') for r in data)
print(blk['name'][:70], 'inst', toti, 'samples', tot)
allst = {s: sum(gi(r, s) for r in data) for s in st}
print('kernel:', ', '.join('%s %.1f%%' % (s[6:], 100.0 * v / tot) for s, v in sorted(allst.items(), key=lambda x: -x[1])[:9]))
for s0 in range(0, len(data), chunk):
  b = data[s0:s0 + chunk]
  sm = sum(gi(r, '# Samples') for r in b); ins = sum(gi(r, 'Instructions Executed') for r in b)
  if sm < 0.01 * tot: continue
  d = {s: sum(gi(r, s) for r in b) for s in st}
  ops = {}
  for r in b:
    t = r[idx['Source']].split(); op = t[0] if t else ''
    if op.startswith('@') and len(t) > 1: op = t[1]
    ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + gi(r, 'Instructions Executed')
  top = ' '.join('%s:%d%%' % (k, 100 * v / max(ins, 1)) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:4])
  print('rows %5d-%5d inst %5.1f%% samples %5.1f%% | %s | %s' % (
      s0, s0 + chunk, 100.0 * ins / toti, 100.0 * sm / tot,
      ', '.join('%s %.0f%%' % (s[6:], 100.0 * v / sm) for s, v in sorted(d.items(), key=lambda x: -x[1])[:5]), top))
