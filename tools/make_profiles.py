"""Turns the raw outputs of tools/gpu_profiles.sh (gpurun_out/) into the committed, judged summaries
under profiles/: launch-list aggregates, per-kernel DRAM traffic (feeds bench.py's roofline.traffic),
and a compact metric table of every `ncu --set full` capture.  usage: make_profiles.py [round]"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
R = sys.argv[1] if len(sys.argv) > 1 else 'r1'
os.makedirs(P, exist_ok=True)


def read_ncu_csv(path):
  rows = list(csv.reader(open(path)))
  h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
  hdr = rows[h]
  return hdr, [r for r in rows[h + 1:] if len(r) == len(hdr)]


def short(name):
  n = name.replace('void ', '').replace('wlseg::', '')
  return n.split('(')[0][:90]


def launch_summary(src, dst):
  hdr, rows = read_ncu_csv(src)
  ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
  agg = defaultdict(lambda: [0, 0.0])
  for r in rows:
    agg[short(r[ki])][0] += 1
    agg[short(r[ki])][1] += float(r[vi].replace(',', ''))
  tot = sum(v[1] for v in agg.values())
  with open(dst, 'w') as fp:
    fp.write(f'# {os.path.basename(src)}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time '
             '(ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n')
    fp.write('#   us_total  launches  share  kernel\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      fp.write(f'{v[1] / 1e3:10.1f} {v[0]:6d} {100 * v[1] / tot:6.2f}%  {k}\n')


def dram_summary(src, dst):
  hdr, rows = read_ncu_csv(src)
  ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
  ii = hdr.index('ID')
  per = defaultdict(dict)
  names = {}
  for r in rows:
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3,
             'msecond': 1e3}.get(u, 1)
    per[r[ii]][r[mi]] = v * scale
    names[r[ii]] = short(r[ki])
  agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
  for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += m.get('dram__bytes_read.sum', 0.0)
    a[2] += m.get('dram__bytes_write.sum', 0.0)
    a[3] += m.get('gpu__time_duration.sum', 0.0)
  out = {}
  with open(dst, 'w') as fp:
    fp.write(f'# {os.path.basename(src)}: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, by kernel\n')
    fp.write('# launches  MB_read/launch  MB_write/launch  us/launch  GB/s(dram)  kernel\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][3]):
      n = a[0]
      fp.write(f'{n:6d} {a[1] / n / 1e6:12.2f} {a[2] / n / 1e6:12.2f} {a[3] / n:10.1f} {(a[1] + a[2]) / max(a[3], 1e-9) / 1e3:10.0f}  {k}\n')
      out[k] = {'launches': n, 'dram_bytes_per_launch': (a[1] + a[2]) / n, 'us_per_launch': a[3] / n}
  return out


def full_summary(rep, dst):
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units = rows[0], rows[1]
  want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
          'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
          'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'smsp__inst_executed.sum',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem']
  idx = [hdr.index(w) for w in want if w in hdr]
  with open(dst, 'w') as fp:
    fp.write(f'# {os.path.basename(rep)}: ncu --set full --clock-control none, selected metrics per captured launch\n')
    for r in rows[2:]:
      fp.write(f'--- launch {r[0]}\n')
      for i in idx:
        fp.write(f'  {hdr[i]:85s} {r[i][:110]} {units[i]}\n')


def main():
  for name in ('bench_eval.json', 'bench_train.json', 'bench_reference.json', 'layers_eval.txt', 'layers_train.txt',
               'eval_classes.json', 'train_classes.json', 'eval_launches.csv', 'train_launches.csv',
               'bench_train_mixed.json', 'bench_vistas_eval.json', 'bench_vistas_train.json', 'timeline_train.txt',
               'timeline_eval.txt'):
    src = os.path.join(G, f'{R}_{name}')
    if os.path.exists(src):
      shutil.copy(src, os.path.join(P, f'{R}_{name}'))
  for w in ('eval', 'train'):
    src = os.path.join(G, f'{R}_{w}_launches.csv')
    if os.path.exists(src):
      launch_summary(src, os.path.join(P, f'{R}_{w}_launches_summary.txt'))
  traffic = {}
  for name, key in ((f'{R}_eval_conv_dram.csv', 'eval'), (f'{R}_train_dram.csv', 'train')):
    src = os.path.join(G, name)
    if os.path.exists(src):
      traffic[key] = dram_summary(src, os.path.join(P, name.replace('.csv', '_summary.txt')))
  if traffic:
    with open(os.path.join(P, f'{R}_traffic.json'), 'w') as fp:
      json.dump(traffic, fp, indent=1, sort_keys=True)
  for f in sorted(os.listdir(G)):
    if f.startswith(R + '_') and f.endswith('.ncu-rep'):
      full_summary(os.path.join(G, f), os.path.join(P, f.replace('.ncu-rep', '_summary.txt')))


if __name__ == '__main__':
  main()
