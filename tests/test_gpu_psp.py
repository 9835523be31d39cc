"""`--psp_module` (code/models/resnet50_extended_model_hierarchical.py:186-207): the four pyramid kernels
against the oracle's restatement of slim.avg_pool2d / tf.image.resize_images and torch autograd of it,
then the whole network with the module switched on (forward, and one training step in the fp32 check
mode) against the oracle graph with `psp=True`.

Tolerances: fp32 kernels 1e-5 of max|ref| (same arithmetic, different summation order); bf16 one output
rounding (1e-2 of max|ref|); network forward 2e-2 (bf16) / 1e-4 (fp32); training step as
tests/test_gpu_train.py (losses 1e-4, gradient cosine >= 0.999, rel-L2 <= 5e-2).
"""

import pytest
import torch

from oracle import losses as olosses
from oracle import network as onet
from oracle import tfops

pytestmark = pytest.mark.gpu


def _rel(got, ref):
  return float((got.double() - ref.double()).abs().max()) / max(float(ref.abs().max()), 1e-30)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize('shape,k', [((2, 12, 13, 256), (6, 6)), ((1, 12, 13, 64), (12, 13)), ((2, 9, 7, 8), (1, 1)),
                                     ((1, 16, 32, 256), (2, 5))])
def test_avgpool_valid_fwd_bwd(cuda, dtype, tol, shape, k):
  from wlseg import ops
  g = torch.Generator().manual_seed(shape[1] * 31 + k[0])
  x = torch.randn(shape, generator=g).to(dtype)
  N, H, W, C = shape
  P, Q = (H - k[0]) // k[0] + 1, (W - k[1]) // k[1] + 1
  y = torch.empty((N, P, Q, C), dtype=dtype, device=cuda)
  ops.avgpool_valid_fwd(x.to(cuda), y, k[0], k[1])
  xr = x.float().requires_grad_(True)
  yr = tfops.avg_pool_valid(xr, k, k)
  assert tuple(yr.shape) == tuple(y.shape)
  assert _rel(y.float().cpu(), yr.detach()) <= tol
  dy = torch.randn(yr.shape, generator=g).to(dtype)
  yr.backward(dy.float())
  # plain and accumulating forms
  dx = torch.empty(shape, dtype=dtype, device=cuda)
  ops.avgpool_valid_bwd(dy.to(cuda), dx, k[0], k[1])
  assert _rel(dx.float().cpu(), xr.grad) <= tol
  base = torch.randn(shape, generator=g).to(dtype)
  dx2 = base.to(cuda)
  ops.avgpool_valid_bwd(dy.to(cuda), dx2, k[0], k[1], accumulate=True)
  assert _rel(dx2.float().cpu(), xr.grad + base.float()) <= tol


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize('src,dst,C', [((1, 1), (12, 13), 256), ((2, 2), (12, 13), 256), ((3, 3), (12, 12), 64),
                                       ((6, 6), (12, 24), 256), ((4, 7), (9, 7), 8), ((5, 5), (5, 5), 16)])
def test_resize_bilinear_fwd_bwd(cuda, dtype, tol, src, dst, C):
  from wlseg import ops
  g = torch.Generator().manual_seed(src[0] * 7 + dst[1])
  N = 2
  x = torch.randn((N, src[0], src[1], C), generator=g).to(dtype)
  # forward into a channel slice of a wider tensor (the PSP concatenation), the rest must stay untouched
  wide = torch.full((N, dst[0], dst[1], 3 * C), 7.0, dtype=dtype, device=cuda)
  ops.resize_bilinear_fwd(x.to(cuda), wide[..., C:2 * C])
  xr = x.float().requires_grad_(True)
  yr = tfops.resize_bilinear(xr, dst[0], dst[1], align_corners=True)
  assert _rel(wide[..., C:2 * C].float().cpu(), yr.detach()) <= tol
  assert bool((wide[..., :C] == 7.0).all()) and bool((wide[..., 2 * C:] == 7.0).all())
  dy = torch.randn((N, dst[0], dst[1], 3 * C), generator=g).to(dtype)
  yr.backward(dy[..., C:2 * C].float())
  dx = torch.empty_like(x, device=cuda)
  ops.resize_bilinear_bwd(dy.to(cuda)[..., C:2 * C], dx)
  assert _rel(dx.float().cpu(), xr.grad) <= tol


def _setup(cuda, dtype, seed, train=False, psp=True, fov=None):
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  tf_params = onet.init_params('cityscapes', seed=seed, randomize_bn=True, tame=True, psp=psp, fov=fov)
  params = network.Params(hier, cuda, psp=psp, fov=fov)
  params.load_tf_dict(tf_params)
  cls = network.TrainNetwork if train else network.Network
  return hier, tf_params, params, cls(params, dtype=dtype)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_forward_with_psp_matches_oracle(cuda, dtype, tol):
  hier, tf_params, params, net = _setup(cuda, dtype, 4)
  assert len(params.specs) == 71
  g = torch.Generator().manual_seed(21)
  images = torch.rand(2, 96, 104, 3, generator=g) * 2 - 1   # 12 x 13 features: the 2- and 3-bin pools drop a column
  out = net.predict(images.to(cuda), want=('decisions',))
  torch.cuda.synchronize()
  ref = onet.Net(tf_params, 'cityscapes', psp=True).forward(images)
  ref_low = torch.cat(ref['lowres_logits'], -1)
  got_low = out['lowres_logits'][..., :hier.total_channels].cpu()
  emax = _rel(got_low, ref_low)
  el2 = float((got_low - ref_low).norm() / ref_low.norm())
  print(f'psp forward {dtype}: low-res logits max-rel {emax:.3e} rel-L2 {el2:.3e}')
  assert emax <= tol and el2 <= tol
  # the module must matter: the same weights without it give different logits
  ref_plain = torch.cat(onet.Net(tf_params, 'cityscapes', psp=False).forward(images)['lowres_logits'], -1)
  assert float((ref_plain - ref_low).norm() / ref_low.norm()) > 10 * tol


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize('fov', [(3, 2), (5, 3)])
def test_forward_with_fov_expansion_matches_oracle(cuda, dtype, tol, fov):
  """--fov_expansion_kernel_size / _rate: the dilated `increase_fov` convolution
  (code/models/resnet50_extended_feature_extractor.py:44-49)."""
  hier, tf_params, params, net = _setup(cuda, dtype, 8, psp=False, fov=fov)
  assert len(params.specs) == 67
  g = torch.Generator().manual_seed(fov[0])
  images = torch.rand(1, 96, 128, 3, generator=g) * 2 - 1
  out = net.predict(images.to(cuda), want=('decisions',))
  torch.cuda.synchronize()
  ref_low = torch.cat(onet.Net(tf_params, 'cityscapes', fov=fov).forward(images)['lowres_logits'], -1)
  got_low = out['lowres_logits'][..., :hier.total_channels].cpu()
  emax, el2 = _rel(got_low, ref_low), float((got_low - ref_low).norm() / ref_low.norm())
  print(f'fov {fov} {dtype}: low-res logits max-rel {emax:.3e} rel-L2 {el2:.3e}')
  assert emax <= tol and el2 <= tol


@pytest.mark.parametrize('psp,fov', [(True, None), (True, (3, 2))])
def test_train_step_with_psp_fp32(cuda, psp, fov):
  hier, tf_params, params, net = _setup(cuda, torch.float32, 6, train=True, psp=psp, fov=fov)
  H, W = 96, 104
  g = torch.Generator().manual_seed(33)
  images = torch.rand(2, H, W, 3, generator=g) * 2 - 1
  labels = {'prolabels_per_pixel': torch.randint(0, 20, (2, H // 8, W // 8), generator=g, dtype=torch.int32)
            .repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()}
  logits = net.forward_train(images.to(cuda))
  losses, dlogits = net.loss_and_grad(logits, {k: v.to(cuda) for k, v in labels.items()}, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  p = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in tf_params.items()}
  onet_ = onet.Net(p, 'cityscapes', training=True, psp=psp, fov=fov)
  rl = olosses.define_losses(onet_.forward(images), labels, 'cityscapes')
  rl['total'].backward()
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  print('losses', losses.cpu().tolist(), want.tolist())
  assert torch.allclose(losses.cpu(), want, rtol=1e-4, atol=1e-5)
  got_all, ref_all, worst = [], [], (1.0, None)
  for s in params.specs:
    ref = p[f'{s.scope}/weights'].grad.permute(3, 0, 1, 2).reshape(-1)
    o = params.w_off[s.scope]
    got = net.ws.grads[o:o + ref.numel()].cpu()
    got_all.append(got)
    ref_all.append(ref)
    if ref.numel() >= 4096:
      c = float(torch.dot(got.double(), ref.double()) / (got.double().norm() * ref.double().norm()))
      if c < worst[0]:
        worst = (c, s.scope)
  ga, ra = torch.cat(got_all).double(), torch.cat(ref_all).double()
  rel = float((ga - ra).norm() / ra.norm())
  print(f'psp train fp32: worst per-tensor cosine {worst[0]:.6f} ({worst[1]}), rel-L2 {rel:.3e}')
  assert worst[0] >= 0.999 and rel <= 5e-2
  for sc in ('feature_extractor/pyramid_module/Conv', 'feature_extractor/pyramid_module/Conv_3',
             'feature_extractor/pyramid_module/Conv_4') + (('feature_extractor/extension/increase_fov',) if fov else ()):
    gr = p[f'{sc}/weights'].grad
    assert float(gr.abs().max()) > 0, sc   # the pyramid branches do receive gradient in the oracle ...
    o = params.w_off[sc]
    got = net.ws.grads[o:o + gr.numel()].cpu().double()
    ref = gr.permute(3, 0, 1, 2).reshape(-1).double()
    c = float(torch.dot(got, ref) / (got.norm() * ref.norm()))
    assert c >= 0.999, (sc, c)               # ... and the same one here


# ---- --upsampling_method hybrid / no (code/models/resnet50_extended_model_hierarchical.py:143-184) ---------------
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize('dataset', ['cityscapes', 'vistas'])
def test_forward_hybrid_upsampling_matches_oracle(cuda, dtype, tol, dataset):
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
  tf_params = onet.init_params(dataset, seed=12, randomize_bn=True, tame=True, upsampling='hybrid')
  params = network.Params(hier, cuda, upsampling='hybrid')
  assert len(params.specs) == 69
  params.load_tf_dict(tf_params)
  back = params.to_tf_dict()   # the transposed-convolution filters survive the layout round trip
  for k in tf_params:
    assert torch.allclose(back[k], tf_params[k].float(), atol=0, rtol=0), k
  net = network.Network(params, dtype=dtype)
  g = torch.Generator().manual_seed(5)
  images = torch.rand(2, 64, 96, 3, generator=g) * 2 - 1
  out = net.predict(images.to(cuda), want=('decisions', 'l1_probabilities'))
  torch.cuda.synchronize()
  ref = onet.Net(tf_params, dataset, upsampling='hybrid').forward(images)
  ref_low = torch.cat(ref['lowres_logits'], -1)
  got_low = out['lowres_logits'][..., :hier.total_channels].cpu()
  emax, el2 = _rel(got_low, ref_low), float((got_low - ref_low).norm() / ref_low.norm())
  print(f'hybrid {dataset} {dtype}: post-upsampler low-res logits max-rel {emax:.3e} rel-L2 {el2:.3e}')
  assert emax <= tol and el2 <= tol
  plain = torch.cat(onet.Net(tf_params, dataset).forward(images)['lowres_logits'], -1)
  assert float((plain - ref_low).norm() / ref_low.norm()) > 10 * tol   # the layer is not a no-op
  assert float((out['decisions'].cpu() != ref['decisions']).float().mean()) <= (1e-3 if dtype == torch.float32 else 0.05)


def test_forward_no_upsampling(cuda):
  """`upsampled = bottom`: probabilities and decisions at the feature resolution."""
  hier, tf_params, _, _ = _setup(cuda, torch.float32, 9, psp=False)
  from wlseg import network
  params = network.Params(hier, cuda, upsampling='no')
  params.load_tf_dict(tf_params)
  net = network.Network(params, dtype=torch.float32)
  g = torch.Generator().manual_seed(6)
  images = torch.rand(1, 64, 96, 3, generator=g) * 2 - 1
  out = net.predict(images.to(cuda), want=('decisions', 'l1_probabilities'))
  ref = onet.Net(tf_params, 'cityscapes', upsampling='no').forward(images)
  assert tuple(out['decisions'].shape) == (1, 8, 12) == tuple(ref['decisions'].shape)
  assert float((out['l1_probabilities'].cpu() - ref['l1_probabilities']).abs().max()) <= 1e-4
  assert float((out['decisions'].cpu() != ref['decisions']).float().mean()) <= 0.02


def test_train_step_hybrid_upsampling_fp32(cuda):
  """One training step with the transposed-convolution upsampler: losses, and the gradients of its filters
  (in TF's [kh, kw, out, in] layout) and biases against autograd of the oracle."""
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  tf_params = onet.init_params('cityscapes', seed=14, randomize_bn=True, tame=True, upsampling='hybrid')
  params = network.Params(hier, cuda, upsampling='hybrid')
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=torch.float32)
  H, W = 64, 96
  g = torch.Generator().manual_seed(41)
  images = torch.rand(2, H, W, 3, generator=g) * 2 - 1
  labels = {'prolabels_per_pixel': torch.randint(0, 20, (2, H // 8, W // 8), generator=g, dtype=torch.int32)
            .repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()}
  logits = net.forward_train(images.to(cuda))
  losses, dlogits = net.loss_and_grad(logits, {k: v.to(cuda) for k, v in labels.items()}, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  p = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in tf_params.items()}
  rl = olosses.define_losses(onet.Net(p, 'cityscapes', training=True, upsampling='hybrid').forward(images), labels,
                             'cityscapes')
  rl['total'].backward()
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  assert torch.allclose(losses.cpu(), want, rtol=1e-4, atol=1e-5)
  got = params.arena_to_tf_dict(net.ws.grads)
  for sc in onet.UPSAMPLING_SCOPES:
    for v in ('weights', 'biases'):
      a, b = got[f'{sc}/{v}'].double().reshape(-1), p[f'{sc}/{v}'].grad.double().reshape(-1)
      cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
      rel = float((a - b).norm() / b.norm())
      print(f'{sc}/{v}: cosine {cos:.6f} rel-L2 {rel:.2e}')
      assert cos >= 0.999 and rel <= 5e-2, (sc, v)
  # and the gradient keeps flowing into the network below the upsampler
  w = 'feature_extractor/base/resnet_v1_50/block3/unit_2/bottleneck_v1/conv2/weights'
  a, b = got[w].double().reshape(-1), p[w].grad.double().reshape(-1)
  assert float(torch.dot(a, b) / (a.norm() * b.norm())) >= 0.999
