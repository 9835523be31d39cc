"""Join two tools/layer_table.py outputs by layer signature: usage cmp_layers.py a.txt b.txt [class-filter]"""
import sys


def load(p):
  d = {}
  for l in open(p).read().splitlines():
    f = l.split()
    if len(f) < 18 or not f[0].isdigit():
      continue
    d[tuple(f[:12])] = (float(f[12]), float(f[13]), float(f[17]))
  return d


a, b = load(sys.argv[1]), load(sys.argv[2])
flt = sys.argv[3] if len(sys.argv) > 3 else ''
ta = tb = 0
for k in sorted(a, key=lambda k: -a[k][0] * a[k][1]):
  if k not in b or flt not in k[11] + k[10]:
    continue
  n, ua, ea = a[k]
  _, ub, eb = b[k]
  ta += n * ua
  tb += n * ub
  print(' '.join('%5s' % x for x in k[:10]), '%-6s %-12s %4.0f x %7.1f -> %7.1f us (%+5.1f%%) eff %.2f -> %.2f' % (
      k[10], k[11], n, ua, ub, 100 * (ub - ua) / ua, ea, eb))
print('total us/step: %.0f -> %.0f' % (ta, tb))
