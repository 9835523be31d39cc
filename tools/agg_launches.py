"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
  if 'Kernel Name' in r:
    h = i
    break
hdr = rows[h]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = defaultdict(lambda: [0, 0.0])
n = 0
for r in rows[h + 1:]:
  if len(r) <= vi:
    continue
  try:
    v = float(r[vi].replace(',', ''))
  except ValueError:
    continue
  name = r[ki].split('(')[0][-70:]
  agg[name][0] += 1
  agg[name][1] += v
  n += 1
tot = sum(v[1] for v in agg.values())
print(f'{n} launches, {tot / 1e6:.3f} ms of kernel time')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
  print(f'{v[1] / 1e3:10.1f} us {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}')
