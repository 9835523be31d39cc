"""Instruction mix (executed warp instructions by opcode) of one kernel from `ncu --page source --csv`. usage: ncu_mix.py source.csv"""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[starts[0]]
si = hdr.index('Source')
ie = hdr.index('# Instructions Executed') if '# Instructions Executed' in hdr else hdr.index('Instructions Executed')
data = [r for r in rows[starts[0] + 1:] if len(r) > si]
cnt = collections.Counter()
tot = 0
for r in data:
  f = r[si].strip().split()
  if not f:
    continue
  op = f[1] if f[0].startswith('@') and len(f) > 1 else f[0]
  n = int(r[ie] or 0)
  cnt[op.split('.')[0]] += n
  tot += n
print('total warp instructions', tot)
for k, v in cnt.most_common(28):
  print(f'{k:12s} {v:10d} {100 * v / tot:5.1f}%')
