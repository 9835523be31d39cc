"""`--norm_layer group` (tf.contrib.layers.group_norm, 32 groups / 1 for the logits layers;
code/models/resnet50_extended_model_hierarchical.py:75-77,314-333): kernels against the oracle's restatement and
its autograd, then the whole network forward and one fp32 training step.

Tolerances: fp32 1e-4 (measured 4e-6 forward, 8e-5 rel-L2 on the whole gradient).  bf16 product path 5e-2 against the
fp32 oracle (measured 2.9e-2): unlike inference batch norm (folded constants, 8e-3), group norm normalises every
layer by the statistics of the stored (bf16-rounded) tensor itself, which - as for train-mode batch norm,
tests/test_gpu_train.py - amplifies the storage rounding through the 66 layers of a random-init network."""

import pytest
import torch

from oracle import losses as olosses
from oracle import network as onet
from oracle import tfops

pytestmark = pytest.mark.gpu


def _rel(a, b):
  return float((a.double() - b.double()).abs().max()) / max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize('N,H,W,C,G,relu,res', [(2, 6, 5, 64, 32, True, False), (3, 4, 7, 256, 32, True, True),
                                                (2, 5, 5, 14, 1, False, False), (1, 8, 8, 768, 96, True, False)])
def test_group_norm_layer_fwd_bwd_fp32(cuda, N, H, W, C, G, relu, res):
  """The group-norm pipeline of network.TrainNetwork (_gn_forward / _gn_backward) on one layer."""
  from wlseg import ops
  g = torch.Generator().manual_seed(C + H)
  z = torch.randn(N, H, W, C, generator=g)
  gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
  r = torch.randn(N, H, W, C, generator=g) if res else None
  da = torch.randn(N, H, W, C, generator=g)
  # oracle
  zr, gr, br = z.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
  rr = None if r is None else r.clone().requires_grad_(True)
  y = tfops.group_norm(zr, gr, br, G, 1e-5)
  if rr is not None:
    y = y + rr
  if relu:
    y = torch.relu(y)
  y.backward(da)
  # product
  hw = H * W
  zd, ad = z.to(cuda), torch.empty(N, H, W, C, device=cuda)
  sums = torch.zeros(2, N, C, dtype=torch.float64, device=cuda)
  for n in range(N):
    ops.bn_stats(zd[n], hw, C, C, sums[0, n], sums[1, n])
  coef = torch.empty(4, N, C, device=cuda)
  ops.gn_finalize(sums[0], sums[1], N, C, G, hw, gamma.to(cuda), beta.to(cuda), 1e-5, coef[0], coef[1], coef[2], coef[3])
  rd = None if r is None else r.to(cuda)
  for n in range(N):
    ops.bn_apply(zd[n], coef[0, n], coef[1, n], None if rd is None else rd[n], ad[n], hw, C, relu)
  assert _rel(ad.cpu(), y.detach()) <= 1e-5
  dad = da.to(cuda)
  yact = ad if (relu and (res or C % 8 != 0)) else None
  part = torch.zeros(2, N, C, dtype=torch.float64, device=cuda)
  for n in range(N):
    ops.bn_bwd_reduce(dad[n], None if yact is None else yact[n], zd[n], coef[2, n], coef[3, n], hw, C, relu, part[0, n],
                      part[1, n], scale=coef[0, n], shift=coef[1, n], pitch=C)
  k = torch.empty(3, N, C, device=cuda)
  dgam, dbet = torch.zeros(C, dtype=torch.float64, device=cuda), torch.zeros(C, dtype=torch.float64, device=cuda)
  ops.gn_bwd_finalize(part[0], part[1], N, C, G, hw, gamma.to(cuda), coef[2], coef[3], k[0], k[1], k[2], dgam, dbet)
  dz = torch.empty_like(zd)
  dres = torch.empty_like(zd) if res else None
  ops.gn_bwd_apply(dad, yact, zd, k[0], k[1], k[2], coef[0], coef[1], N, hw, C, relu, dz, dres)
  assert _rel(dz.cpu(), zr.grad) <= 1e-4
  assert _rel(dgam.cpu(), gr.grad) <= 1e-4 and _rel(dbet.cpu(), br.grad) <= 1e-4
  if res:
    assert _rel(dres.cpu(), rr.grad) <= 1e-6


def _setup(cuda, dtype, seed):
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  tf_params = onet.init_params('cityscapes', seed=seed, randomize_bn=True, tame=True, norm='group')
  params = network.Params(hier, cuda, norm='group')
  params.load_tf_dict(tf_params)
  back = params.to_tf_dict()
  assert set(back) == set(tf_params) and not any('BatchNorm' in k for k in back)
  return hier, tf_params, params, network.TrainNetwork(params, dtype=dtype)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-4), (torch.bfloat16, 5e-2)])
def test_forward_group_norm_matches_oracle(cuda, dtype, tol):
  hier, tf_params, params, net = _setup(cuda, dtype, 21)
  g = torch.Generator().manual_seed(8)
  images = torch.rand(2, 64, 96, 3, generator=g) * 2 - 1
  out = net.predict(images.to(cuda), want=('decisions',))
  torch.cuda.synchronize()
  ref = onet.Net(tf_params, 'cityscapes', norm='group').forward(images)
  ref_low = torch.cat(ref['lowres_logits'], -1)
  got_low = out['lowres_logits'][..., :hier.total_channels].cpu()
  emax, el2 = _rel(got_low, ref_low), float((got_low - ref_low).norm() / ref_low.norm())
  print(f'group norm forward {dtype}: low-res logits max-rel {emax:.3e} rel-L2 {el2:.3e}')
  assert emax <= tol and el2 <= tol


def test_train_step_group_norm_fp32(cuda):
  hier, tf_params, params, net = _setup(cuda, torch.float32, 23)
  H, W = 64, 96
  g = torch.Generator().manual_seed(77)
  images = torch.rand(2, H, W, 3, generator=g) * 2 - 1
  labels = {'prolabels_per_pixel': torch.randint(0, 20, (2, H // 8, W // 8), generator=g, dtype=torch.int32)
            .repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()}
  logits = net.forward_train(images.to(cuda))
  losses, dlogits = net.loss_and_grad(logits, {k: v.to(cuda) for k, v in labels.items()}, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  p = {k: v.clone().requires_grad_(True) for k, v in tf_params.items()}
  rl = olosses.define_losses(onet.Net(p, 'cityscapes', training=True, norm='group').forward(images), labels, 'cityscapes')
  rl['total'].backward()
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  print('losses', losses.cpu().tolist(), want.tolist())
  assert torch.allclose(losses.cpu(), want, rtol=1e-4, atol=1e-5)
  got = params.arena_to_tf_dict(net.ws.grads)
  worst, ga, ra = (1.0, None), [], []
  for k, v in got.items():
    a, b = v.double().reshape(-1), p[k].grad.double().reshape(-1)
    if 'weights' in k:
      b = p[k].grad.double().reshape(-1)
    ga.append(a)
    ra.append(b)
    if a.numel() >= 4096:
      c = float(torch.dot(a, b) / (a.norm() * b.norm()))
      if c < worst[0]:
        worst = (c, k)
  ga, ra = torch.cat(ga), torch.cat(ra)
  rel = float((ga - ra).norm() / ra.norm())
  print(f'group norm train fp32: worst per-tensor cosine {worst[0]:.6f} ({worst[1]}), rel-L2 {rel:.3e}')
  assert worst[0] >= 0.999 and rel <= 5e-2
