"""Debug helper (GPU + CPU oracle): per-layer comparison of the bf16 training forward against the
oracle graph with explicit bf16 storage roundings.  Not a test."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch
from oracle import network as onet
from wlseg import arch, hierarchy, network, problem_defs

dev = torch.device('cuda:0')
dataset = 'cityscapes'
hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
tf_params = onet.init_params(dataset, seed=11, randomize_bn=True, tame=True)
H, W = int(os.environ.get('H', 64)), int(os.environ.get('W', 96))
g = torch.Generator().manual_seed(18)
images = torch.rand(2, H, W, 3, generator=g) * 2 - 1
params = network.Params(hier, dev)
params.load_tf_dict(tf_params)
net = network.TrainNetwork(params, dtype=torch.bfloat16)
logits = net.forward_train(images.to(dev))
torch.cuda.synchronize()
o = onet.Net(tf_params, dataset, training=True, storage='bf16')
o.record_layers = True
with torch.no_grad():
  low = o.lowres_logits(images)
d = 256
s0 = 'adaptation_module/l1_features/conv1'
for s in params.specs:
  if s.scope not in net.tape:
    continue
  rec = net.tape[s.scope]
  zo, ao = o.layer_taps[s.scope]
  zg, ag = rec.z.float().cpu(), rec.a.float().cpu()
  if s.scope == s0:
    zg, ag = zg[..., :d], ag[..., :d]
  ez = float((zg - zo).norm() / (zo.norm() + 1e-20))
  ea = float((ag - ao).norm() / (ao.norm() + 1e-20))
  nz = float((zg != zo).float().mean())
  print(f'{s.scope[-58:]:58s} z {ez:.3e} (mismatching elems {nz:.4f}) a {ea:.3e}')
