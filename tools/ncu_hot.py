"""Top stall sites of one kernel from `ncu --page source --csv` output (SASS view).
usage: ncu_hot.py source.csv [top] [block]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = starts[blk]
rows = rows[:starts[blk + 1] - 1] if blk + 1 < len(starts) else rows
hdr = rows[hi]
si, ss = hdr.index('Source'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[hi + 1:] if len(r) > ss]
tot = sum(int(r[ss] or 0) for r in data)
print('total samples', tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
order = sorted(range(len(data)), key=lambda i: -int(data[i][ss] or 0))[:top]
for i in sorted(order):
  r = data[i]
  st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
  print(f'{i:5d} {int(r[ss]):7d} {100.0 * int(r[ss]) / tot:5.1f}%  {r[si].strip()[:90]:90s} {st}')
