"""Per-layer comparison of the gradient arena between the fused and the separate BN-backward reduction."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch
from oracle import network as onet
from wlseg import hierarchy, network, problem_defs
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_train import _labels

cuda = torch.device('cuda:0')
hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
tf_params = onet.init_params('cityscapes', seed=17, randomize_bn=True, tame=True)
g = torch.Generator().manual_seed(23)
H, W = int(os.environ.get('DBG_H', 96)), int(os.environ.get('DBG_W', 136))
images = (torch.rand(2, H, W, 3, generator=g) * 2 - 1).to(cuda)
labels = {k: v.to(cuda) for k, v in _labels('cityscapes', 2, 0, 0, H, W, 20).items()}
params = network.Params(hier, cuda)
params.load_tf_dict(tf_params)
net = network.TrainNetwork(params, dtype=torch.bfloat16)
logits = net.forward_train(images)
losses, dlogits = net.loss_and_grad(logits, labels, H, W)
n = params.n_chan_pad
runs = []
for flag in (True, False):
  net.bnb_fuse = flag
  net.ws.stat[2 * n:].zero_()
  runs.append(net.backward(dlogits).cpu().clone())
g1, g0 = runs
p = params
for s in p.specs:
  o = p.w_off[s.scope]; k = s.K * s.R * s.S * s.C
  a, b = g1[o:o + k], g0[o:o + k]
  e = float((a - b).abs().max() / (b.abs().max() + 1e-30))
  c = p.c_off[s.scope]
  ga, gb = g1[p.n_conv_pad + c:p.n_conv_pad + c + s.K], g0[p.n_conv_pad + c:p.n_conv_pad + c + s.K]
  ba, bb = g1[p.n_conv_pad + n + c:p.n_conv_pad + n + c + s.K], g0[p.n_conv_pad + n + c:p.n_conv_pad + n + c + s.K]
  eg = float((ga - gb).abs().max() / (gb.abs().max() + 1e-30)); eb = float((ba - bb).abs().max() / (bb.abs().max() + 1e-30))
  flag = ' <<<' if max(e, eg, eb) > 1e-4 else ''
  print(f'{s.scope[-45:]:>45s} K{s.K:5d} R{s.R} s{s.stride} d{s.dilation}  dw {e:.2e}  dgamma {eg:.2e}  dbeta {eb:.2e}{flag}')
