import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch
from wlseg import hierarchy, network, ops, problem_defs, synthetic, trainer as wtrainer
dev = torch.device('cuda:0')
hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
params = network.Params(hier, dev); params.init_random(0)
class S:
  momentum, use_nesterov, optimizer, regularization_weight = 0.9, False, 'SGDM', 0.00017
  batch_norm_decay, distribute, ema_decay = 0.9, False, 0.0
tr = wtrainer.Trainer(params, S)
src = synthetic.SyntheticInputs(hier.num_classes, dev)
f, l = src.train_batch(4, 0, 0, 768, 768)
l = {k: v for k, v in l.items() if v is not None}
print('image dtype', f['proimages'].dtype, {k: (v.dtype, tuple(v.shape)) for k, v in l.items()})
def timeit(tag, n=5):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for i in range(n): tr.step(f, l, 0.01)
  torch.cuda.synchronize(); print(tag, (time.perf_counter() - t0) / n * 1e3, 'ms/step', 'graphs', len(tr._graphs), 'mem GB', torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
timeit('warm (2 eager + capture)', 3)
timeit('replay')
tr.net.profile = []
timeit('eager profiled', 3)
tr.net.profile = None
timeit('replay after eager')
timeit('replay after eager 2')
img_h = (torch.rand((4, 768, 768, 3)) * 2 - 1).pin_memory()
lab_h = torch.randint(0, 20, (4, 768, 768), dtype=torch.int32).pin_memory()
for rep in range(3):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for i in range(5):
    fi = {'proimages': img_h.to(dev, non_blocking=True)}
    li = {'prolabels_per_pixel': lab_h.to(dev, non_blocking=True)}
    tr.step(fi, li, 0.01)
  torch.cuda.synchronize(); print('host-fed', rep, (time.perf_counter() - t0) / 5 * 1e3, 'ms/step', 'graphs', len(tr._graphs), torch.cuda.memory_reserved() / 1e9)
