"""Per-layer table of the convolution kernels measured in situ (CUDA events around every launch of a
whole forward / training step): count, mean us, TFLOP/s, algorithmic GB/s and the roofline bound
(max of flops / tensor peak and bytes / HBM peak).  usage: layer_table.py [eval|train] [H W N]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402
from wlseg import hierarchy, network, problem_defs, synthetic, trainer as wtrainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'eval'
H, W, N = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else \
    ((1024, 2048, 4) if mode == 'eval' else (768, 768, 4))
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) \
    else {'hbm_gbs': 6650.0, 'bf16_tflops_sustained': 1400.0}
dev = torch.device('cuda:0')
hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
params = network.Params(hier, dev)
params.init_random(0)
src = synthetic.SyntheticInputs(hier.num_classes, dev)
recs = []
if mode == 'eval':
  net = network.Network(params)
  f, l = src.eval_batch(N, H, W)
  for i in range(5):
    if i == 2:
      net.profile = recs
    net.predict(f['proimages'])
  reps = 3
else:
  class S:
    momentum, use_nesterov, optimizer, regularization_weight = 0.9, False, 'SGDM', 0.00017
    batch_norm_decay, distribute, ema_decay = 0.9, False, 0.0
  tr = wtrainer.Trainer(params, S, use_graph=False)
  f, l = src.train_batch(N, 0, 0, H, W)
  for i in range(5):
    if i == 2:
      tr.net.profile = recs
    tr.step(f, {k: v for k, v in l.items() if v is not None}, 0.01)
  reps = 3
torch.cuda.synchronize()
agg = {}
for r in recs:
  a = agg.setdefault(r['sig'], {'n': 0, 'ms': 0.0, 'flops': r['flops'], 'bytes': r['bytes'], 'cls': r['cls']})
  a['n'] += 1
  a['ms'] += r['e0'].elapsed_time(r['e1'])
bw = {k: v for k, v in agg.items() if len(k) < 11}
agg = {k: v for k, v in agg.items() if len(k) >= 11}
tot = sum(a['ms'] for a in agg.values()) / reps
print(f'{mode} {N}x{H}x{W}: conv kernels {tot:.3f} ms/step')
for sig, a in bw.items():
  us = 1e3 * a['ms'] / a['n']
  print(f'  bandwidth kernel {sig[0]}: {a["n"] / reps:.1f}/step, {us:.1f} us, algorithmic {a["bytes"] / 1e6:.1f} MB -> '
        f'{a["bytes"] / us / 1e3:.0f} GB/s = {a["bytes"] / us / 1e3 / peaks["hbm_gbs"]:.3f} of the measured HBM peak')
print('   N    H    W    C    K R s d res bn kind    cls          n/step   us    TF/s   GB/s  bound_us  eff')
for sig, a in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
  us = 1e3 * a['ms'] / a['n']
  bound = max(a['flops'] / (peaks['bf16_tflops_sustained'] * 1e12), a['bytes'] / (peaks['hbm_gbs'] * 1e9)) * 1e6
  print('%4d %4d %4d %4d %4d %d %d %d %3d %2d %-6s %-12s %5.1f %7.1f %6.0f %6.0f %8.1f %5.2f' % (
      sig[:8] + (int(sig[8]), int(sig[9]), sig[10], a['cls'], a['n'] / reps, us, a['flops'] / us / 1e6,
                 a['bytes'] / us / 1e3, bound, bound / us)))
