"""Parity at the BASELINE.json shapes themselves (not only at toy sizes): the CUDA path against the CPU oracle on the
same seeded inputs and weights, at 1024x2048 (configs[1]), 4 x 768 x 768 (configs[2]) and Vistas 1080x1920
(configs[4]).  The oracle needs 2-15 s of host time per case.

Tolerances (north star): low-res logits within 2e-2 (max-rel and rel-L2) of the fp32 oracle on the bf16 product path;
decisions: disagreement rate bounded near what is measured (near-ties flip under bf16); the three losses within 2e-2
relative; the confusion matrix BIT-EXACT against the numpy and the plain-C oracle on the same decisions.
"""

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import metrics as ometrics
from oracle import network as onet

pytestmark = pytest.mark.gpu


def _setup(cuda, dataset, seed):
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
  tf_params = onet.init_params(dataset, seed=seed, randomize_bn=True, tame=True)
  params = network.Params(hier, cuda)
  params.load_tf_dict(tf_params)
  return hier, tf_params, params


def _errors(got, ref):
  return float((got - ref).abs().max() / ref.abs().max()), float((got - ref).norm() / ref.norm())


@pytest.mark.parametrize('dataset,H,W', [('cityscapes', 1024, 2048), ('vistas', 1080, 1920)])
def test_eval_forward_full_size_matches_oracle(cuda, dataset, H, W):
  """One full-size evaluation image: logits, the four decision maps and the confusion matrix."""
  from wlseg import network, ops
  torch.set_num_threads(max(1, torch.get_num_threads()))
  hier, tf_params, params = _setup(cuda, dataset, seed=21)
  net = network.Network(params, dtype=torch.bfloat16)
  ncls = hier.num_classes
  g = torch.Generator().manual_seed(H + W)
  images = torch.rand(1, H, W, 3, generator=g) * 2 - 1
  labels = torch.randint(0, ncls, (1, H, W), generator=g, dtype=torch.int32)
  out = net.predict(images.to(cuda), want=('decisions', 'l1_decisions', 'l2_vehicle_decisions', 'l2_human_decisions'))
  cm = torch.zeros(ncls, ncls, dtype=torch.int64, device=cuda)
  ops.confmat_accumulate(labels.to(cuda), out['decisions'], ncls, cm)
  torch.cuda.synchronize()
  with torch.no_grad():
    ref = onet.Net(tf_params, dataset).forward(images)
  ref_low = torch.cat(ref['lowres_logits'], -1)
  got_low = out['lowres_logits'][..., :hier.total_channels].cpu()
  assert got_low.shape == ref_low.shape == (1, (H + 7) // 8, (W + 7) // 8, hier.total_channels)
  emax, el2 = _errors(got_low, ref_low)
  decs = out['decisions'].cpu()
  dis = float((decs != ref['decisions']).float().mean())
  dis1 = float((out['l1_decisions'].cpu() != ref['l1_decisions']).float().mean())
  print(f'{dataset} 1x{H}x{W}: low-res logits max-rel {emax:.3e} rel-L2 {el2:.3e}; decision disagreement {dis:.4f} '
        f'(l1 head {dis1:.4f})')
  assert emax <= 2e-2 and el2 <= 2e-2
  assert dis <= 0.02 and dis1 <= 0.02
  # the integer tail: bit-exact on the SAME decisions, against both restatements
  assert np.array_equal(cm.cpu().numpy(), ometrics.confusion_matrix(labels.numpy(), decs.numpy(), ncls))
  assert np.array_equal(cm.cpu().numpy(), ometrics.confusion_matrix_c(labels.numpy(), decs.numpy(), ncls))
  # and the head kernel alone, fed the ORACLE's low-res logits: decisions identical except at exact near-ties
  low = torch.zeros((1,) + tuple(ref_low.shape[1:3]) + (hier.logits_pitch,), dtype=torch.float32)
  low[..., :hier.total_channels] = ref_low
  hd = net.head(low.to(cuda), H, W, ('decisions',))['decisions'].cpu()
  assert float((hd != ref['decisions']).float().mean()) <= 1e-4


def _conditioned_params(dataset, seed, res_gamma=0.2):
  """Random init with the residual-branch gammas scaled down.  A train-mode BN ResNet at plain random init is a
  chaotic map: perturbations grow ~1.08x per layer (x300 end to end - the oracle with bf16 storage roundings and the
  fp32 oracle decorrelate to 0.8 rel-L2 on the logits, profiles/r2_train_parity_probe.txt), so NO bf16 implementation
  can be compared with an fp32 one there.  Small residual gammas give the conditioning of a trained network."""
  tf_params = onet.init_params(dataset, seed=seed, randomize_bn=True, tame=True)
  for k in tf_params:
    if k.endswith('conv3/BatchNorm/gamma') and 'bottleneck' in k:
      tf_params[k] = tf_params[k] * res_gamma
  return tf_params


def _oracle_train_step(tf_params, dataset, images, labels, storage):
  p = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in tf_params.items()}
  pred = onet.Net(p, dataset, training=True, storage=storage).forward(images)
  rl = olosses.define_losses(pred, labels, dataset)
  rl['total'].backward()
  low = torch.cat(pred['lowres_logits'], -1).detach()
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  return low, want, {k: v.grad for k, v in p.items() if v.requires_grad}


def _grad_cosines(params, grads, ref):
  got_all, ref_all, worst = [], [], (1.0, None)
  for s in params.specs:
    r = ref[f'{s.scope}/weights'].permute(3, 0, 1, 2).reshape(-1)
    o = params.w_off[s.scope]
    gt = grads[o:o + r.numel()]
    got_all.append(gt)
    ref_all.append(r)
    if r.numel() >= 4096:
      c = float(torch.dot(gt.double(), r.double()) / (gt.double().norm() * r.double().norm()))
      if c < worst[0]:
        worst = (c, s.scope)
  ga, ra = torch.cat(got_all).double(), torch.cat(ref_all).double()
  return float(torch.dot(ga, ra) / (ga.norm() * ra.norm())), worst


def test_train_step_4x768_matches_oracle(cuda):
  """BASELINE configs[2]: one 4 x 768 x 768 strong-label training step of the bf16 PRODUCT path (tcgen05
  convolutions incl. the CTA-pair forms, fused batch statistics, masked residual gradients, fused loss, backward)
  end to end against the oracle's autograd - (a) the fp32 oracle and (b) the oracle with the product's storage
  roundings made explicit (oracle/network.py storage='bf16').  Measured on B200 (tools/train_parity_probe.py,
  profiles/r2_train_parity_probe.txt): losses 3e-5..1.2e-4; logits rel-L2 6.1e-2 vs (a), 3.7e-2 vs (b) - and (b) vs (a)
  is itself 6.0e-2, i.e. the product is as close to the fp32 oracle as a bf16-storage implementation can be; gradient
  cosine 0.919 global / 0.894 worst tensor vs (a), 0.957 / 0.946 vs (b).  The asserts bound each figure at about
  1.5x the measured gap, and tie the product's distance to (a) to the oracle's own (b)-vs-(a) distance."""
  from wlseg import network
  dataset, N, H, W = 'cityscapes', 4, 768, 768
  from wlseg import hierarchy, problem_defs
  hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
  tf_params = _conditioned_params(dataset, seed=31)
  params = network.Params(hier, cuda)
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=torch.bfloat16)
  g = torch.Generator().manual_seed(77)
  images = torch.rand(N, H, W, 3, generator=g) * 2 - 1
  labels = {'prolabels_per_pixel': torch.randint(0, 20, (N, H // 32, W // 32), generator=g, dtype=torch.int32)
            .repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()}
  logits = net.forward_train(images.to(cuda))
  losses, dlogits = net.loss_and_grad(logits, {k: v.to(cuda) for k, v in labels.items()}, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  got_low = logits[..., :hier.total_channels].cpu()
  got_losses = losses.cpu()
  grads = net.ws.grads.cpu()
  del net, logits, dlogits
  torch.cuda.empty_cache()

  low32, want32, g32 = _oracle_train_step(tf_params, dataset, images, labels, 'fp32')
  low16, want16, g16 = _oracle_train_step(tf_params, dataset, images, labels, 'bf16')
  e32 = _errors(got_low, low32)[1]
  e16 = _errors(got_low, low16)[1]
  e_or = _errors(low16, low32)[1]
  lerr32 = float(((got_losses - want32).abs() / want32.abs()).max())
  lerr16 = float(((got_losses - want16).abs() / want16.abs()).max())
  cos32, worst32 = _grad_cosines(params, grads, g32)
  cos16, worst16 = _grad_cosines(params, grads, g16)
  ga = torch.cat([g16[f'{s.scope}/weights'].reshape(-1) for s in params.specs]).double()
  gb = torch.cat([g32[f'{s.scope}/weights'].reshape(-1) for s in params.specs]).double()
  cos_or = float(torch.dot(ga, gb) / (ga.norm() * gb.norm()))
  print(f'4x768x768 train step: losses {got_losses.tolist()} vs fp32 oracle {want32.tolist()} (max rel {lerr32:.2e}; vs '
        f'bf16-storage oracle {lerr16:.2e}); logits rel-L2 {e32:.3e} vs fp32 oracle, {e16:.3e} vs bf16-storage oracle '
        f'(oracle bf16-storage vs oracle fp32: {e_or:.3e}); gradient cosine global {cos32:.4f} / worst {worst32[0]:.4f} '
        f'({worst32[1]}) vs fp32 oracle, {cos16:.4f} / {worst16[0]:.4f} vs bf16-storage oracle (oracle vs oracle {cos_or:.4f})')
  assert lerr32 <= 2e-3 and lerr16 <= 2e-3            # north star: loss within 2e-2 in bf16
  assert e16 <= TRAIN_LOGITS_REL_L2_VS_BF16_ORACLE
  assert e32 <= 1.5 * e_or + 1e-3                     # no further from fp32 than the storage roundings explain
  assert cos32 >= TRAIN_GRAD_COS_GLOBAL and worst32[0] >= TRAIN_GRAD_COS_WORST
  assert cos16 >= TRAIN_GRAD_COS_GLOBAL_VS_BF16_ORACLE
  assert cos32 >= cos_or - 0.05


# ~1.5x the gaps measured on B200 (profiles/r2_train_parity_probe.txt, RES_GAMMA=0.2 rows)
TRAIN_LOGITS_REL_L2_VS_BF16_ORACLE = 6e-2
TRAIN_GRAD_COS_GLOBAL = 0.88
TRAIN_GRAD_COS_WORST = 0.84
TRAIN_GRAD_COS_GLOBAL_VS_BF16_ORACLE = 0.93


def test_confmat_full_size_vs_oracle(cuda):
  """BASELINE configs[1] step size (4 x 1024 x 2048 pixels) against the numpy and plain-C oracles (bit-exact), with
  segmentation-like runs AND uniformly random pairs."""
  from wlseg import ops
  rng = np.random.default_rng(9)
  n = 4 * 1024 * 2048
  lab = np.repeat(np.repeat(rng.integers(0, 20, size=(4, 32, 64), dtype=np.int32), 32, axis=1), 32, axis=2).reshape(-1)
  dec = rng.integers(0, 20, size=n, dtype=np.int32)
  cm = torch.zeros(20, 20, dtype=torch.int64, device=cuda)
  ops.confmat_accumulate(torch.from_numpy(lab).to(cuda), torch.from_numpy(dec).to(cuda), 20, cm)
  torch.cuda.synchronize()
  got = cm.cpu().numpy()
  assert got.sum() == n
  assert np.array_equal(got, ometrics.confusion_matrix_c(lab, dec, 20))
  assert np.array_equal(got, ometrics.confusion_matrix(lab, dec, 20))


def test_batch_mean_iou_product_function_on_device_path(cuda):
  """define_metrics.mean_iou (code/estimator/define_metrics.py:5-20) as the product computes it: confusion matrix from
  the CUDA histogram kernel on device tensors -> estimator.mean_iou_from_cm, against the oracle's batch_mean_iou."""
  from wlseg import estimator, ops
  rng = np.random.default_rng(4)
  for C, shape in ((20, (4, 96, 128)), (66, (2, 135, 240)), (5, (1, 7, 9))):
    lab = rng.integers(0, C, size=shape, dtype=np.int32)
    dec = np.where(rng.random(shape) < 0.6, lab, rng.integers(0, C, size=shape, dtype=np.int32)).astype(np.int32)
    if C == 66:
      dec[lab == 3] = 4          # a class with zero intersection
      lab[lab == 7] = 8          # a class absent from the labels (union > 0 through the decisions only)
    cm = torch.zeros(C, C, dtype=torch.int64, device=cuda)
    ops.confmat_accumulate(torch.from_numpy(lab).to(cuda), torch.from_numpy(dec).to(cuda), C, cm)
    torch.cuda.synchronize()
    got = estimator.mean_iou_from_cm(cm.cpu().numpy(), C)
    want = float(ometrics.batch_mean_iou(lab, dec, C))
    assert abs(got - want) <= 1e-6 * max(1.0, abs(want)), (C, got, want)
  # known answer on a hand-made matrix: IoUs 1/2, 1/3, 0 (union 1), 0 (union 0 -> 0 / 1e-9) -> mean over ALL 4 classes
  cm4 = np.array([[1, 1, 0, 0], [0, 1, 0, 0], [0, 1, 0, 0], [0, 0, 0, 0]])
  assert abs(estimator.mean_iou_from_cm(cm4, 4) - (0.5 + 1 / 3) / 4) < 1e-6


def test_predict_config0_512x1024_matches_oracle(cuda):
  """BASELINE configs[0] - the reference's own CPU-runnable case: predict.py's forward, batch 1, one 512 x 1024
  Cityscapes-shaped image, random init - on the bf16 product path against the fp32 CPU oracle on the same input and
  weights: the keys predict.py requests (decisions, l1 and l2_vehicle probability maps at full resolution).
  Low-res logits rel-L2 <= 2e-2 (measured 7-9e-3 at every other shape); decisions: disagreement <= 2 % (Cityscapes heads
  measure <= 0.15 %); probability maps: mean absolute difference <= 1e-2."""
  from wlseg import network
  H, W = 512, 1024
  hier, tf_params, params = _setup(cuda, 'cityscapes', seed=22)
  net = network.Network(params, dtype=torch.bfloat16)
  g = torch.Generator().manual_seed(H + W)
  images = torch.rand(1, H, W, 3, generator=g) * 2 - 1
  out = net.predict(images.to(cuda), want=('decisions', 'l1_probabilities', 'l2_vehicle_probabilities'))
  torch.cuda.synchronize()
  with torch.no_grad():
    ref = onet.Net(tf_params, 'cityscapes').forward(images)
  ref_low = torch.cat(ref['lowres_logits'], -1)
  got_low = out['lowres_logits'][..., :hier.total_channels].cpu()
  emax, el2 = _errors(got_low, ref_low)
  decs = out['decisions'].cpu()
  assert tuple(decs.shape) == (1, H, W) and decs.dtype == torch.int32
  dis = float((decs != ref['decisions']).float().mean())
  dp1 = float((out['l1_probabilities'].cpu() - ref['l1_probabilities']).abs().mean())
  dpv = float((out['l2_vehicle_probabilities'].cpu() - ref['l2_vehicle_probabilities']).abs().mean())
  print(f'configs[0] 1x{H}x{W}: low-res logits max-rel {emax:.3e} rel-L2 {el2:.3e}; decision disagreement {dis:.4f}; '
        f'mean |dp| l1 {dp1:.2e} l2_vehicle {dpv:.2e}')
  assert tuple(out['l1_probabilities'].shape) == (1, H, W, 14) and tuple(out['l2_vehicle_probabilities'].shape) == (1, H, W, 7)
  assert el2 <= 2e-2 and dis <= 2e-2 and dp1 <= 1e-2 and dpv <= 1e-2
