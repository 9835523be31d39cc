"""GPU probe (not a test): the bf16 PRODUCT training forward, free running, against the oracle with the product's
storage roundings (oracle/network.py storage='bf16') and against the fp32 oracle, layer by layer: conv output z and
activation a.  Shows whether the end-to-end gap accumulates smoothly (rounding noise of a deep BN network) or jumps
at one layer (a bug).   RES_GAMMA=0.2 python tools/debug_train_layers.py 2x64x96"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402

from oracle import network as onet  # noqa: E402
from wlseg import hierarchy, network, problem_defs  # noqa: E402

dev = torch.device('cuda:0')
dataset = 'cityscapes'
hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
N, H, W = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '2x64x96').split('x'))
fac = float(os.environ.get('RES_GAMMA', '1.0'))
tf_params = onet.init_params(dataset, seed=11, randomize_bn=True, tame=True)
for k in tf_params:
  if k.endswith('conv3/BatchNorm/gamma') and 'bottleneck' in k:
    tf_params[k] = tf_params[k] * fac
g = torch.Generator().manual_seed(18)
images = torch.rand(N, H, W, 3, generator=g) * 2 - 1
params = network.Params(hier, dev)
params.load_tf_dict(tf_params)
net = network.TrainNetwork(params, dtype=torch.bfloat16)
net.keep = True
net.forward_train(images.to(dev))
torch.cuda.synchronize()
taps = {}
for storage in ('fp32', 'bf16'):
  with torch.no_grad():
    o = onet.Net(tf_params, dataset, training=True, storage=storage)
    o.record_layers = True
    o.layer_taps = {}
    o.forward(images)
  taps[storage] = o.layer_taps


def rel(a, b):
  return float((a - b).norm() / b.norm().clamp_min(1e-30))


print(f'[res_gamma {fac}] {N}x{H}x{W}: product(bf16) vs oracle[bf16 storage] | product vs oracle[fp32] | oracle[bf16] vs oracle[fp32]')
for spec in params.specs:
  if spec.scope not in net.tape or spec.scope not in taps['bf16']:
    continue
  rec = net.tape[spec.scope]
  K = rec.nch
  z, a = rec.z.float().cpu()[..., :K], rec.a.float().cpu()[..., :K]
  zb, ab = taps['bf16'][spec.scope]
  zf, af = taps['fp32'][spec.scope]
  if z.shape != zb.shape:
    continue
  print(f'{spec.scope[-44:]:44s} z {rel(z, zb):.2e} a {rel(a, ab):.2e} | z {rel(z, zf):.2e} a {rel(a, af):.2e} | z {rel(zb, zf):.2e} a {rel(ab, af):.2e}')
