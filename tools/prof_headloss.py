"""The head kernel at the evaluation size (4 x 1024 x 2048) and the strong / mixed loss kernels at the training size
(768 x 768), a few launches each, for ncu:  python tools/prof_headloss.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402
from wlseg import hierarchy, network, ops, problem_defs, synthetic  # noqa: E402

dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
hs = hier.as_struct()
src = synthetic.SyntheticInputs(20, dev)


def timed(fn, name, nbytes):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 1e3 / reps
  print(f'{name}: {us:8.1f} us  ({nbytes / us / 1e3:7.1f} GB/s algorithmic)')


# ---- evaluation head: low-res logits -> decisions (+ confusion matrix)
N, H, W = 4, 1024, 2048
low = torch.randn(N, H // 8, W // 8, hier.logits_pitch, device=dev) * 3
labels = src.strong_labels(N, H, W)
net = network.Network.__new__(network.Network)
net.hier, net.hstruct, net.dev = hier, hs, dev
cm = torch.zeros(20, 20, dtype=torch.int64, device=dev)


def head_cm():
  d = net.head(low, H, W, ('decisions',))['decisions']
  ops.confmat_accumulate(labels, d, 20, cm)


timed(head_cm, 'head + confmat 4x1024x2048', N * H * W * 8 + low.numel() * 4)
timed(lambda: ops.head_confmat(hs, low, H, W, labels, 20, cm), 'fused head_confmat 4x1024x2048', N * H * W * 4 + low.numel() * 4)
dbuf = torch.empty((N, H, W), dtype=torch.int32, device=dev)
timed(lambda: ops.head_confmat(hs, low, H, W, None, 20, None, None, None, dbuf), 'head_confmat, decisions only', N * H * W * 4 + low.numel() * 4)

# ---- training losses at 768 x 768
H, W = 768, 768
for name, (ns, nb, ni), compact in (('strong 4', (4, 0, 0), False), ('mixed 4+8+4 dense', (4, 8, 4), False),
                                    ('mixed 4+8+4 lists', (4, 8, 4), True)):
  B = ns + nb + ni
  logits = torch.randn(B, H // 8, W // 8, hier.logits_pitch, device=dev) * 2
  f, l = src.train_batch(ns, nb, ni, H, W, compact=compact)
  dl = torch.zeros_like(logits)
  sums = torch.zeros(3, dtype=torch.float64, device=dev)
  counts = torch.zeros(3, dtype=torch.float64, device=dev)
  if compact:
    fn = lambda: ops.loss_fwd_bwd_lists(hs, logits, H, W, l['prolabels_per_pixel'], l.get('bbox_coords'), l.get('bbox_cids'),
                                        l.get('image_vectors'), sums, counts, dl)
    nbytes = ns * H * W * 4 + 2 * logits.numel() * 4
  else:
    fn = lambda: ops.loss_fwd_bwd(hs, logits, H, W, l['prolabels_per_pixel'], l.get('prolabels_per_bbox'),
                                  l.get('prolabels_per_image'), sums, counts, dl)
    nbytes = ns * H * W * 4 + (nb + ni) * H * W * 60 + 2 * logits.numel() * 4
  timed(fn, f'loss {name} at 768x768', nbytes)
