"""SASS evidence per kernel of libwlseg.so (cuobjdump runs without a GPU): tensor-core, TMEM, TMA and cluster
mnemonics per entry point.  usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200', 'wlseg', 'lib', 'libwlseg.so')
KEYS = ['UTCHMMA.2CTA', 'UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'SYNCS', 'UCGABAR',
        'MUFU.RCP', 'MEMBAR', 'ATOMS', 'ATOMG', 'RED.', 'HMMA']
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
funcs = collections.OrderedDict()
cur = None
for line in out.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    cur = funcs.setdefault(m.group(1), collections.Counter())
    continue
  if cur is None:
    continue
  m = re.search(r'/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
  if not m:
    continue
  op = m.group(1)
  cur['_instr'] += 1
  for k in KEYS:
    if op.startswith(k) or (k.endswith('.') and op.startswith(k[:-1] + '.')):
      cur[k] += 1
      break
print(f'# {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels, cuobjdump -sass mnemonic counts (sm_100a)')
print(f'# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG/UTMASTG/UTMAREDG/UTMAPF = TMA load / store / '
      f'reduce-add / L2 prefetch, UTCBAR = tcgen05.commit, UCGABAR = barrier.cluster, SYNCS = mbarrier ops, MUFU.RCP = a '
      f'runtime integer or float division')
print(f'{"instr":>7} ' + ' '.join(f'{k:>12}' for k in KEYS) + '  kernel')
tot = collections.Counter()
for name, c in funcs.items():
  tot.update(c)
  short = re.sub(r'\(.*', '', demangle(name))
  print(f'{c["_instr"]:7d} ' + ' '.join(f'{c[k]:12d}' for k in KEYS) + f'  {short[:110]}')
print(f'{tot["_instr"]:7d} ' + ' '.join(f'{tot[k]:12d}' for k in KEYS) + '  TOTAL')
