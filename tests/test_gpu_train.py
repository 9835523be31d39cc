"""Parity of the TRAINING path (forward with batch statistics -> hierarchical masked loss ->
backward through the whole network) with the oracle's autograd on the same
inputs and weights.

A train-mode batch-norm ResNet at random init is a strongly amplifying map: on these shapes the
oracle itself, evaluated in fp32 and in fp64, disagrees with itself by 2.1e-2 (relative L2 of the
whole gradient; worst per-tensor cosine 0.9996), and rounding only the INPUT image to bf16 changes
its logits by 30 % and drops the gradient cosine to 0.53 (measured with this file's `_oracle_step`).
Hence:
  fp32 check mode (direct fp32 convolutions) vs the fp32 oracle: losses 1e-4 relative; gradient
      cosine >= 0.999 for every tensor >= 4096 elements, global relative L2 <= 5e-2 (the oracle's
      own fp32 noise floor is 2.1e-2)
  bf16 product path (tcgen05 convolutions, bf16 storage) vs the SAME oracle graph with the storage
      roundings made explicit (`Net(storage='bf16')`): logits 2e-2 relative L2, losses 2e-2
      relative (north star), gradient cosine >= 0.99 per large tensor and over the whole arena
"""

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import network as onet
from oracle import weak_labels as oweak

pytestmark = pytest.mark.gpu


def _oracle_step(tf_params, dataset, images, labels, storage='fp32'):
  params = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in tf_params.items()}
  net = onet.Net(params, dataset, training=True, storage=storage)
  pred = net.forward(images)
  losses = olosses.define_losses(pred, labels, dataset)
  losses['total'].backward()
  grads = {k: v.grad for k, v in params.items() if v.requires_grad}
  return losses, grads, net.new_moving, pred


def _labels(dataset, n_strong, n_bbox, n_image, H, W, seed):
  g = torch.Generator().manual_seed(seed)
  ncls = 20 if dataset == 'cityscapes' else 66
  lab = {'prolabels_per_pixel': torch.randint(0, ncls, (n_strong, H // 8, W // 8), generator=g, dtype=torch.int32)
         .repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()}
  if n_bbox:
    boxes = []
    for _ in range(n_bbox):
      k = int(torch.randint(1, 6, (1,), generator=g))
      cids = torch.randint(0, 14, (k,), generator=g).tolist()
      c = torch.rand(k, 4, generator=g)
      boxes.append([(cid, float(min(a, b)), float(max(a, b)), float(min(cc, d)), float(max(cc, d)))
                    for cid, (a, b, cc, d) in zip(cids, c.tolist())])
    lab['prolabels_per_bbox'] = torch.stack([torch.from_numpy(oweak.bbox_labels(b, H, W)) for b in boxes])
  if n_image:
    lab['prolabels_per_image'] = torch.stack([torch.from_numpy(oweak.image_labels([int(torch.randint(0, 14, (1,), generator=g))], H, W))
                                              for _ in range(n_image)])
  return lab


def _cos(a, b):
  a, b = a.double().reshape(-1), b.double().reshape(-1)
  return float(torch.dot(a, b) / (a.norm() * b.norm()))


def _compare(net, params, grads, tag):
  """-> (min cosine over tensors with >= 4096 elements, global cosine, global rel-L2)."""
  got_all, ref_all = [], []
  worst = (1.0, None)
  for s in params.specs:
    ref = grads[f'{s.scope}/weights'].permute(3, 0, 1, 2).reshape(-1)  # HWIO -> KRSC
    o = params.w_off[s.scope]
    got = net.ws.grads[o:o + ref.numel()].cpu()
    got_all.append(got)
    ref_all.append(ref)
    c = s_off = params.c_off[s.scope]
    del c
    for name, base in (('gamma', params.n_conv_pad), ('beta', params.n_conv_pad + params.n_chan_pad)):
      r = grads[f'{s.scope}/BatchNorm/{name}']
      gg = net.ws.grads[base + s_off:base + s_off + s.K].cpu()
      got_all.append(gg)
      ref_all.append(r)
    if ref.numel() >= 4096:
      cos = _cos(got, ref)
      if cos < worst[0]:
        worst = (cos, s.scope)
  ga, ra = torch.cat(got_all).double(), torch.cat(ref_all).double()
  gcos = _cos(ga, ra)
  rel = float((ga - ra).norm() / ra.norm())
  print(f'{tag}: worst per-tensor cosine {worst[0]:.6f} ({worst[1]}), global cosine {gcos:.6f}, rel-L2 {rel:.3e}')
  return worst[0], gcos, rel


def _run(cuda, dataset, dtype, n_strong, n_bbox, n_image, H, W, seed, storage='fp32'):
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
  tf_params = onet.init_params(dataset, seed=seed, randomize_bn=True, tame=True)
  params = network.Params(hier, cuda)
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=dtype)
  g = torch.Generator().manual_seed(seed + 7)
  N = n_strong + n_bbox + n_image
  images = torch.rand(N, H, W, 3, generator=g) * 2 - 1
  labels = _labels(dataset, n_strong, n_bbox, n_image, H, W, seed + 9)
  dev_labels = {k: v.to(cuda) for k, v in labels.items()}
  logits = net.forward_train(images.to(cuda))
  losses, dlogits = net.loss_and_grad(logits, dev_labels, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  ref_losses, ref_grads, ref_moving, ref_pred = _oracle_step(tf_params, dataset, images, labels, storage)
  return hier, params, net, logits, losses.cpu(), ref_losses, ref_grads, ref_moving, ref_pred


def test_train_step_fp32_check_mode(cuda):
  hier, params, net, logits, losses, rl, rg, rmov, rpred = _run(cuda, 'cityscapes', torch.float32, 2, 0, 0, 64, 64, 3)
  ref_low = torch.cat(rpred['lowres_logits'], -1)
  got_low = logits[..., :hier.total_channels].cpu()
  assert float((got_low - ref_low.detach()).abs().max()) <= 1e-3 * float(ref_low.abs().max())
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  print('losses', losses.tolist(), want.tolist())
  assert torch.allclose(losses, want, rtol=1e-4, atol=1e-5)
  worst, gcos, rel = _compare(net, params, rg, 'fp32')
  assert worst >= 0.999 and gcos >= 0.999 and rel <= 5e-2
  # moving statistics after the step (decay 0.9, unbiased variance)
  for scope in ('feature_extractor/base/resnet_v1_50/conv1', 'feature_extractor/extension/decrease_fdims',
                'softmax_classifier/l1_logits'):
    mm = params.moving_mean(scope).cpu()
    mv = params.moving_var(scope).cpu()
    assert torch.allclose(mm, rmov[f'{scope}/BatchNorm/moving_mean'].detach(), rtol=1e-3, atol=1e-4), scope
    assert torch.allclose(mv, rmov[f'{scope}/BatchNorm/moving_variance'].detach(), rtol=1e-3, atol=1e-4), scope


def test_train_step_fp32_weak_labels(cuda):
  """Mixed strong + bbox + image-level batch (BASELINE configs[3] shape class, small)."""
  hier, params, net, logits, losses, rl, rg, _, _ = _run(cuda, 'cityscapes', torch.float32, 1, 1, 1, 48, 64, 5)
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  print('losses', losses.tolist(), want.tolist())
  assert torch.allclose(losses, want, rtol=1e-4, atol=1e-5)
  worst, gcos, rel = _compare(net, params, rg, 'fp32 weak')
  assert worst >= 0.999 and gcos >= 0.999 and rel <= 5e-2


def _ref_conv_geom(x, w_krsc, geom):
  """Plain fp32 convolution with the layer's explicit geometry (leading pad, output size, stride,
  dilation): x NHWC, w KRSC -> NPQK."""
  import torch.nn.functional as F
  pad, out_hw, stride, dilation = geom
  _, H, W, _ = x.shape
  _, R, S, _ = w_krsc.shape
  P, Q = out_hw
  pb = max(0, (P - 1) * stride + (R - 1) * dilation + 1 - H - pad[0])
  pr = max(0, (Q - 1) * stride + (S - 1) * dilation + 1 - W - pad[1])
  xp = F.pad(x.permute(0, 3, 1, 2), (pad[1], pr, pad[0], pb))
  y = F.conv2d(xp, w_krsc.permute(0, 3, 1, 2), stride=stride, dilation=dilation)
  return y[:, :, :P, :Q].permute(0, 2, 3, 1)


def _close(got, ref, tol, what):
  err = float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)
  assert err <= tol, f'{what}: max-rel {err:.3e} > {tol}'
  return err


@pytest.mark.parametrize('dataset,psp,fov', [('cityscapes', False, None), ('vistas', False, None),
                                             ('cityscapes', True, (3, 2))])
def test_train_step_bf16_layerwise(cuda, dataset, psp, fov):
  """bf16 product path, every layer checked IN SITU against a plain fp32 restatement fed with the
  pipeline's own (bf16) inputs: forward conv, batch statistics, BN+residual+ReLU, BN backward,
  filter gradient, data gradient (+ fused gradient fan-in).  End-to-end comparison with the oracle is
  meaningless in bf16 for this graph (module docstring); layer-local parity plus the fp32 end-to-end
  test above pins both the kernels and their wiring.
  Tolerances: one bf16 rounding of the output (1e-2 of max|ref|) for bf16 tensors; 2e-3 for fp32
  outputs (dw, dgamma, dbeta, statistics)."""
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
  tf_params = onet.init_params(dataset, seed=11, randomize_bn=True, tame=True, psp=psp, fov=fov)
  params = network.Params(hier, cuda, psp=psp, fov=fov)   # + the pyramid module and increase_fov layers
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=torch.bfloat16)
  net.keep = True
  H, W = 64, 96
  g = torch.Generator().manual_seed(18)
  images = torch.rand(2, H, W, 3, generator=g) * 2 - 1
  labels = _labels(dataset, 2, 0, 0, H, W, 20)
  logits = net.forward_train(images.to(cuda))
  losses, dlogits = net.loss_and_grad(logits, {k: v.to(cuda) for k, v in labels.items()}, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  ws, n_checked = net.ws, 0
  worst = {}
  for spec in params.specs:
    if spec.scope not in net.tape:
      continue  # the 2nd / 3rd adaptation conv1 live inside the merged 256 -> 768 layer
    rec = net.tape[spec.scope]
    K, off = rec.nch, params.c_off[spec.scope]
    x = rec.x.float().cpu().requires_grad_(True)
    w = rec.w.float().cpu().requires_grad_(True)
    z_ref = _ref_conv_geom(x, w, rec.geom)
    zq = rec.z.float().cpu()
    e = {}
    e['z'] = _close(zq, z_ref.detach(), 1e-2, f'{spec.scope} conv output')
    n = z_ref.shape[0] * z_ref.shape[1] * z_ref.shape[2]
    # batch statistics are DEFINED over the stored (bf16-rounded) conv output, the tensor the normalisation
    # is applied to; with few samples per channel (the 1-bin pyramid branch has
    # n = batch size) the statistics of the unrounded z_ref differ by the rounding itself
    mean_ref = zq.double().mean((0, 1, 2))
    var_ref = zq.double().var((0, 1, 2), unbiased=False)
    mean = ws.view(ws.bn, 2, off, K).cpu()
    invstd = ws.view(ws.bn, 3, off, K).cpu()
    assert torch.allclose(mean.double(), mean_ref, rtol=2e-3, atol=2e-3 * float(var_ref.sqrt().max())), spec.scope
    assert torch.allclose(invstd.double(), (var_ref + 1e-5).rsqrt(), rtol=2e-3), spec.scope
    gamma, beta = params.gamma(spec.scope, K).cpu(), params.beta(spec.scope, K).cpu()
    zhat = (zq - mean) * invstd
    a_ref = zhat * gamma + beta
    if rec.res is not None:
      a_ref = a_ref + rec.res.float().cpu()
    if rec.relu:
      a_ref = torch.relu(a_ref)
    a = rec.a.float().cpu()
    e['a'] = _close(a, a_ref, 1e-2, f'{spec.scope} activation')
    # ---- backward, teacher forced with the pipeline's own incoming gradient
    da = rec.da.float().cpu()
    gg = da * (a > 0).float() if rec.relu else da
    dbeta_ref = gg.double().sum((0, 1, 2))
    dgamma_ref = (gg.double() * zhat.double()).sum((0, 1, 2))
    gb = net.ws.grads[params.n_conv_pad + params.n_chan_pad + off:params.n_conv_pad + params.n_chan_pad + off + K].cpu()
    gm = net.ws.grads[params.n_conv_pad + off:params.n_conv_pad + off + K].cpu()
    e['dbeta'] = _close(gb.double(), dbeta_ref, 2e-3, f'{spec.scope} dbeta')
    e['dgamma'] = _close(gm.double(), dgamma_ref, 2e-3, f'{spec.scope} dgamma')
    dz_ref = (gamma * invstd) * (gg - (dbeta_ref / n).float() - zhat * (dgamma_ref / n).float())
    dz = rec.dz.float().cpu()[..., :K]
    e['dz'] = _close(dz, dz_ref, 1e-2, f'{spec.scope} dz')
    z_ref.backward(dz)
    o = params.w_off[spec.scope]
    if rec.kind == 'root_packed':
      # filter gradient checked in the ORIGINAL 7x7/2 geometry (validates the pack + gather)
      img = images.to(torch.bfloat16).float().requires_grad_(False)
      w7 = params.w32(spec.scope).to(torch.bfloat16).float().cpu().requires_grad_(True)
      z7 = _ref_conv_geom(img, w7, ((3, 3), rec.geom[1], 2, 1))
      z7.backward(dz)
      dw_ref = w7.grad
    else:
      dw_ref = w.grad
    dw = net.ws.grads[o:o + dw_ref.numel()].view(dw_ref.shape).cpu()
    e['dw'] = _close(dw, dw_ref, 2e-3, f'{spec.scope} dw')
    if rec.dx is not None:
      dx_ref = x.grad
      if rec.dx_add is not None:
        dx_ref = dx_ref + rec.dx_add.float().cpu()
      e['dx'] = _close(rec.dx.float().cpu(), dx_ref, 1e-2, f'{spec.scope} dx')
    for k, v in e.items():
      worst[k] = max(worst.get(k, 0.0), v)
    n_checked += 1
  print(f'{dataset}: {n_checked} layers checked in situ; worst max-rel errors {worst}')
  assert n_checked == len(params.specs) - 2


def test_trainer_graph_replay_equals_eager(cuda):
  """The product path replays the whole step as one CUDA graph: starting from the same weights, five
  optimizer steps (2 eager + capture + replays) must give the losses and weights of five eagerly
  launched steps (fp64/fp32 atomics make the two runs differ in the last bits only)."""
  from wlseg import hierarchy, network, problem_defs, trainer as wtrainer

  class S:
    momentum, use_nesterov, optimizer, regularization_weight = 0.9, False, 'SGDM', 0.00017
    batch_norm_decay, distribute, ema_decay = 0.9, False, 0.9

  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  g = torch.Generator().manual_seed(5)
  batches = []
  for _ in range(2):
    img = (torch.rand(2, 64, 96, 3, generator=g) * 2 - 1).to(cuda)
    lab = {'prolabels_per_pixel': torch.randint(0, 20, (2, 64, 96), generator=g, dtype=torch.int32).to(cuda)}
    batches.append((img, lab))
  out = []
  for mode in (False, False, True):
    params = network.Params(hier, cuda)
    params.init_random(3)
    tr = wtrainer.Trainer(params, S, use_graph=mode)
    losses = []
    for i in range(5):
      img, lab = batches[i % 2]
      losses.append(tr.step({'proimages': img}, lab, 0.01 if i < 3 else 0.005).cpu().clone())
    torch.cuda.synchronize()
    assert (len(tr._graphs) == 1) == mode
    out.append((torch.stack(losses), params.master.cpu().clone(), params.moving.cpu().clone(),
                tr.ws.ema_shadow.cpu().clone()))
  # Not bit-exact: fp32 / fp64 atomics (BN statistics, split-K wgrad, loss scatter) commit in a different
  # order from run to run and a train-mode BN ResNet at random init amplifies that (see the file header).
  # The yardstick is therefore the difference between two EAGER runs of the same five steps.
  for e1, e2, gr, name in zip(out[0], out[1], out[2], ('losses', 'weights', 'moving statistics', 'ema shadows')):
    assert torch.isfinite(gr).all(), name
    scale = float(e1.abs().max())
    noise = float((e1 - e2).abs().max()) / scale
    err = float((e1 - gr).abs().max()) / scale
    print(f'{name}: eager-vs-eager {noise:.2e}, graph-vs-eager {err:.2e}')
    assert err <= max(4.0 * noise, 1e-5), f'{name}: graph replay differs from eager by {err:.2e} (run-to-run noise {noise:.2e})'


def test_premasked_residual_gradient_equals_unmasked_wiring(cuda):
  """The ReLU bit masks of the bottleneck outputs (wlseg_bn_apply_mask -> wlseg_conv2d_fprop_masked: the dgrad
  epilogue multiplies the finished gradient by the ReLU derivative, the BN backward skips the activation read and
  the shortcut-gradient write) against the unmasked wiring, as two backward passes over ONE forward tape: bf16
  rounding and masking commute, so every tensor is bit-identical up to the commit order of the fp64 / split-K
  atomics - the yardstick is a third pass with the unmasked wiring again."""
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  tf_params = onet.init_params('cityscapes', seed=13, randomize_bn=True, tame=True)
  g = torch.Generator().manual_seed(21)
  images = (torch.rand(2, 96, 128, 3, generator=g) * 2 - 1).to(cuda)
  labels = {k: v.to(cuda) for k, v in _labels('cityscapes', 2, 0, 0, 96, 128, 20).items()}
  params = network.Params(hier, cuda)
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=torch.bfloat16)
  assert net.premask
  logits = net.forward_train(images)
  assert sum(1 for r in net.tape.values() if getattr(r, 'mask', None) is not None) == 16   # ResNet-50's bottleneck outputs
  # the masks are the sign bits of the stored activations
  rec = net.tape['feature_extractor/base/resnet_v1_50/block2/unit_2/bottleneck_v1/conv3']
  bits = torch.from_numpy(np.unpackbits(rec.mask.cpu().numpy(), axis=1, bitorder='little')).bool()
  assert torch.equal(bits, (rec.a.float().cpu() > 0).reshape(bits.shape))
  losses, dlogits = net.loss_and_grad(logits, labels, 96, 128)
  n = params.n_chan_pad
  runs = []
  for flag in (True, False, False):
    net.premask = flag
    net.ws.stat[2 * n:].zero_()          # dgamma / dbeta accumulators (forward_train zeroes them once per step)
    runs.append(net.backward(dlogits).cpu().clone())
  torch.cuda.synchronize()
  g1, g0, g0b = runs
  scale = float(g0.abs().max())
  noise = float((g0 - g0b).abs().max()) / scale
  err = float((g1 - g0).abs().max()) / scale
  cos = float(torch.dot(g1.double(), g0.double()) / (g1.double().norm() * g0.double().norm()))
  print(f'masked vs unmasked gradient arena: max-rel {err:.2e} (unmasked run-to-run {noise:.2e}), cosine {cos:.10f}')
  assert err <= max(4.0 * noise, 1e-5) and cos >= 0.9999999


def test_fused_bn_backward_reduction_equals_separate_passes(cuda):
  """WLSEG_BNB_FUSE wiring (TrainNetwork.bnb_fuse: inside a bottleneck unit the dgrad of conv3 / conv2 runs as
  wlseg_conv2d_fprop_bnbwd for the BN of conv2 / conv1, whose bn_bwd_reduce pass is never launched) against the
  separate passes, as backward passes over ONE forward tape.
  The masked gradients are bit-identical (tests/test_gpu_conv.py::test_fprop_bnbwd_...), but the fused sums are fp32
  partials in another order than bn_reduce_kernel's: dgamma / dbeta differ by ~1e-7, which flips a few bf16 roundings
  of dz, and a train-mode BN ResNet amplifies any perturbation on its way down (~x300 end to end at plain random init,
  tests/test_gpu_baseline_shapes.py::_conditioned_params).  Hence: (a) 32 launches fewer; (b) the FIRST fused unit
  of the backward pass (block4/unit_3, nothing amplified yet) is tight - conv2's dgamma / dbeta 1e-5 of their
  maximum, its filter gradient 1e-3; (c) on the conditioned network the whole arena agrees to cosine >= 0.9999."""
  from wlseg import hierarchy, network, ops, problem_defs
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  # the conditioning of tests/test_gpu_baseline_shapes.py::_conditioned_params: residual-branch gammas x 0.2
  tf_params = onet.init_params('cityscapes', seed=17, randomize_bn=True, tame=True)
  for k in tf_params:
    if k.endswith('conv3/BatchNorm/gamma') and 'bottleneck' in k:
      tf_params[k] = tf_params[k] * 0.2
  g = torch.Generator().manual_seed(23)
  H, W = 192, 264
  images = (torch.rand(2, H, W, 3, generator=g) * 2 - 1).to(cuda)
  labels = {k: v.to(cuda) for k, v in _labels('cityscapes', 2, 0, 0, H, W, 20).items()}
  params = network.Params(hier, cuda)
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=torch.bfloat16)
  assert net.bnb_fuse
  logits = net.forward_train(images)
  losses, dlogits = net.loss_and_grad(logits, labels, H, W)
  n = params.n_chan_pad
  runs, launches = [], []
  for flag in (True, False, False):
    net.bnb_fuse = flag
    net.ws.stat[2 * n:].zero_()          # dgamma / dbeta accumulators (forward_train zeroes them once per step)
    l0 = ops.launches
    runs.append(net.backward(dlogits).cpu().clone())
    launches.append(ops.launches - l0)
  torch.cuda.synchronize()
  assert launches[1] - launches[0] == 32, f'expected 32 fewer launches in the fused form, got {launches}'
  g1, g0, g0b = runs

  def rel(scope, what):
    sp = params.by_scope[scope]
    if what == 'dw':
      o, k = params.w_off[scope], sp.K * sp.R * sp.S * sp.C
    else:
      o, k = params.n_conv_pad + (n if what == 'dbeta' else 0) + params.c_off[scope], sp.K
    a, b = g1[o:o + k], g0[o:o + k]
    return float((a - b).abs().max() / b.abs().max())

  u = 'feature_extractor/base/resnet_v1_50/block4/unit_3/bottleneck_v1'
  first = {w: rel(f'{u}/conv2', w) for w in ('dgamma', 'dbeta', 'dw')}
  second = {w: rel(f'{u}/conv1', w) for w in ('dgamma', 'dbeta', 'dw')}
  scale = float(g0.abs().max())
  noise = float((g0 - g0b).abs().max()) / scale
  err = float((g1 - g0).abs().max()) / scale
  cos = float(torch.dot(g1.double(), g0.double()) / (g1.double().norm() * g0.double().norm()))
  print(f'fused vs separate BN backward reduction: first fused layer {first}, next {second}; gradient arena max-rel {err:.2e} '
        f'(separate run-to-run {noise:.2e}), cosine {cos:.8f}')
  assert first['dgamma'] <= 1e-5 and first['dbeta'] <= 1e-5 and first['dw'] <= 1e-3
  assert cos >= 0.9995   # measured 0.99996 on B200
