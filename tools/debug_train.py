"""Debug helper (GPU): per-layer comparison of the bf16 training forward/backward against the fp32
check mode on the same inputs.  Not a test; prints the first layers whose error is large."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch
from oracle import network as onet
from wlseg import hierarchy, network, problem_defs

dev = torch.device('cuda:0')
dataset = 'cityscapes'
hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
tf_params = onet.init_params(dataset, seed=11, randomize_bn=True, tame=True)
H, W = int(os.environ.get('H', 64)), int(os.environ.get('W', 96))
g = torch.Generator().manual_seed(18)
images = (torch.rand(2, H, W, 3, generator=g) * 2 - 1).to(dev)
labels = {'prolabels_per_pixel': torch.randint(0, 20, (2, H, W), generator=g, dtype=torch.int32).to(dev)}
nets = {}
for name, dt in (('fp32', torch.float32), ('bf16', torch.bfloat16)):
  params = network.Params(hier, dev)
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=dt)
  logits = net.forward_train(images)
  losses, dlogits = net.loss_and_grad(logits, labels, H, W)
  grads = net.backward(dlogits).clone()
  torch.cuda.synchronize()
  nets[name] = (net, params, logits, losses.clone(), grads)
a, b = nets['fp32'], nets['bf16']
print('losses fp32', a[3].tolist(), 'bf16', b[3].tolist())
for s in a[1].specs:
  ra, rb = a[0].tape[s.scope], b[0].tape[s.scope]
  za, zb = ra.z.float(), rb.z.float()
  aa, ab = ra.a.float(), rb.a.float()
  ez = float((za - zb).norm() / (za.norm() + 1e-20))
  ea = float((aa - ab).norm() / (aa.norm() + 1e-20))
  o = a[1].w_off[s.scope]
  n = s.K * s.R * s.S * s.C
  ga, gb = a[4][o:o + n], b[4][o:o + n]
  cos = float(torch.nn.functional.cosine_similarity(ga, gb, dim=0))
  flag = ' <<<' if (ez > 0.05 or ea > 0.05 or cos < 0.98) else ''
  print(f'{s.scope[-60:]:60s} z {ez:.3e} a {ea:.3e} wgrad-cos {cos:.5f} |g| {float(ga.norm()):.3e}{flag}')
