"""The CUDA kernels (through the C ABI) against vectors produced by RUNNING THE REFERENCE'S OWN PYTHON
(tests/golden/reference_run.npz, written by tests/golden/make_reference_fixtures.py from /root/reference/code over the
TensorFlow-1.12 API emulation in tests/golden/tf_shim).  No oracle in between: reference output vs kernel output.

Tolerances: integer work (decisions, rasterised labels up to the fp32 division, remapped ids, nearest resize) bit-exact;
fp32 kernels 1e-4 relative (north star fp32 check mode), gradient additionally cosine >= 0.99999.
"""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_run.npz')
LOSS_CASES = ['losses_cityscapes_mixed', 'losses_cityscapes_strong', 'losses_vistas_mixed', 'losses_vistas_strong']


@pytest.fixture(scope='module')
def gold():
  return np.load(GOLD)


def _hier(dataset):
  from wlseg import hierarchy, problem_defs
  return hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])


def _packed_logits(gold, tag, hier):
  low = [torch.from_numpy(gold[f'{tag}/lowres_{k}_logits']) for k in ('l1', 'l2_vehicle', 'l2_human')]
  B, h, w = low[0].shape[:3]
  logits = torch.zeros(B, h, w, hier.logits_pitch)
  logits[..., :hier.total_channels] = torch.cat(low, -1)
  return logits


@pytest.mark.parametrize('tag', LOSS_CASES)
@pytest.mark.parametrize('from_boxes', [False, True, 'lists'])
def test_loss_kernel_equals_the_reference_run(cuda, gold, tag, from_boxes):
  """`define_losses` (define_losses_hierarchical.py:14-224) run by the reference: the three losses and d total /
  d low-res logits, against wlseg_loss_fwd_bwd + wlseg_loss_finalize on the same logits and labels.  from_boxes: the
  bbox labels are rasterised on the device from the (class, box) lists (wlseg_rasterize_bbox_labels) instead of
  being uploaded dense, the image-level ones tiled on the device; 'lists': the lists go straight into the loss kernel
  (wlseg_loss_fwd_bwd_lists), which expands them per pixel in registers - no dense weak label exists at all."""
  from wlseg import ops
  n_pp, n_pb, n_pi, h, w = (int(x) for x in gold[f'{tag}/counts'])
  if from_boxes and n_pb + n_pi == 0:
    pytest.skip('strong-only case has no weak labels')
  H, W = 8 * h, 8 * w
  hier = _hier(str(gold[f'{tag}/dataset']))
  hs = hier.as_struct()
  logits = _packed_logits(gold, tag, hier).to(cuda)
  strong = torch.from_numpy(gold[f'{tag}/prolabels_per_pixel']).to(cuda)
  bbox = image = None
  coords = cids = None
  if n_pb:
    if from_boxes:
      mb = max(len(gold[f'{tag}/bbox{i}_cids']) for i in range(n_pb))
      coords = torch.zeros(n_pb, mb, 4)
      cids = torch.full((n_pb, mb), -1, dtype=torch.int32)
      for i in range(n_pb):
        k = len(gold[f'{tag}/bbox{i}_cids'])
        coords[i, :k] = torch.from_numpy(gold[f'{tag}/bbox{i}_coords'])
        cids[i, :k] = torch.from_numpy(gold[f'{tag}/bbox{i}_cids'])
      coords, cids = coords.to(cuda), cids.to(cuda)
      bbox = ops.rasterize_bbox_labels(coords, cids, H, W)
      assert torch.equal(bbox.cpu(), torch.from_numpy(gold[f'{tag}/prolabels_per_bbox']))   # vs _generate_rla, bit-exact
    else:
      bbox = torch.from_numpy(gold[f'{tag}/prolabels_per_bbox']).to(cuda)
  if n_pi:
    vec = torch.from_numpy(gold[f'{tag}/prolabels_per_image_vectors'])
    image = ops.tile_image_labels(vec.to(cuda), H, W) if from_boxes else vec[:, None, None, :].expand(n_pi, H, W, 15).contiguous().to(cuda)
  dl = torch.zeros_like(logits)
  sums = torch.zeros(3, dtype=torch.float64, device=cuda)
  counts = torch.zeros(3, dtype=torch.float64, device=cuda)
  out = torch.zeros(4, device=cuda)
  if from_boxes == 'lists':
    vecd = torch.from_numpy(gold[f'{tag}/prolabels_per_image_vectors']).to(cuda) if n_pi else None
    if hier.head_widths != (14, 7, 3):
      with pytest.raises(ops.WlsegError):      # Vistas heads: the caller rasterises (network.loss_and_grad does)
        ops.loss_fwd_bwd_lists(hs, logits, H, W, strong, coords, cids, vecd, sums, counts, dl)
      return
    ops.loss_fwd_bwd_lists(hs, logits, H, W, strong, coords, cids, vecd, sums, counts, dl)
  else:
    ops.loss_fwd_bwd(hs, logits, H, W, strong, bbox, image, sums, counts, dl)
  ops.loss_finalize(hs, sums, counts, 0.1, 1.0, dl, out)
  torch.cuda.synchronize()
  got = out.cpu().tolist()
  ref = [float(gold[f'{tag}/loss_{k}']) for k in ('l1_segmentation', 'l2_vehicle_segmentation', 'l2_human_segmentation')]
  seg = float(gold[f'{tag}/loss_total']) - float(gold[f'{tag}/loss_regularization'])
  for g, r in zip(got, ref + [seg]):
    assert abs(g - r) <= 1e-4 * max(1.0, abs(r)), (got, ref, seg)
  g = dl.cpu()[..., :hier.total_channels]
  gr = torch.cat([torch.from_numpy(gold[f'{tag}/grad_lowres_{k}_logits']) for k in ('l1', 'l2_vehicle', 'l2_human')], -1)
  assert float((g - gr).abs().max()) <= 1e-4 * float(gr.abs().max()) + 1e-9
  assert float((g * gr).sum() / (g.norm() * gr.norm() + 1e-30)) >= 0.99999


@pytest.mark.parametrize('tag', LOSS_CASES)
def test_head_l1_decisions_equal_the_reference_run(cuda, gold, tag):
  """tf.image.resize_images(align_corners) -> softmax -> argmax as the reference's model code calls them
  (resnet50_extended_model_hierarchical.py:84-93) vs wlseg_head_fwd on the same low-res logits: identical ids."""
  from wlseg import network
  n_pp, n_pb, n_pi, h, w = (int(x) for x in gold[f'{tag}/counts'])
  hier = _hier(str(gold[f'{tag}/dataset']))
  net = network.Network.__new__(network.Network)
  net.hier, net.hstruct, net.dev = hier, hier.as_struct(), cuda
  got = net.head(_packed_logits(gold, tag, hier).to(cuda), 8 * h, 8 * w, ('l1_decisions',))
  assert np.array_equal(got['l1_decisions'].cpu().numpy(), gold[f'{tag}/l1_decisions'])


def test_rasteriser_equals_generate_rla(cuda, gold):
  """input_subset_bboxes_v2.py:74-98 run by the reference vs wlseg_rasterize_bbox_labels (bit-exact, unknown mids
  skipped)."""
  from wlseg import ops
  h, w = (int(x) for x in gold['rla/size'])
  got = ops.rasterize_bbox_labels(torch.from_numpy(gold['rla/coords'])[None].to(cuda), torch.from_numpy(gold['rla/cids'])[None].to(cuda), h, w)
  assert np.array_equal(got.cpu().numpy()[0], gold['rla/out'])


def test_confmat_lut_equals_the_reference_remap(cuda, gold):
  """_map_predictions_to_new_cids (:490-528) run by the reference, then a histogram of its output, vs the LUT fused
  into wlseg_confmat_accumulate."""
  from wlseg import estimator as west
  from wlseg import ops
  for tag, n_old in (('remap', 5), ('remap_cs', 20)):
    decs = torch.from_numpy(gold[f'{tag}/decisions'])
    lut = torch.tensor(west._replacevoids([int(x) for x in gold[f'{tag}/map']]), dtype=torch.int32)
    n_new = int(lut.max()) + 1
    labels = torch.from_numpy(gold[f'{tag}/out_decisions']).to(torch.int32)   # "ground truth" = the reference's remapped ids
    cm = torch.zeros(n_new, n_new, dtype=torch.int64, device=cuda)
    ops.confmat_accumulate(labels.to(cuda), decs.to(cuda), n_new, cm, lut=lut.to(cuda))
    cm = cm.cpu()
    assert int(cm.sum()) == decs.numel() and int(torch.diagonal(cm).sum()) == decs.numel()   # every pixel on the diagonal


@pytest.mark.parametrize('tag', ['resize_up', 'resize_down'])
def test_resize_predictions_equals_the_reference_run(cuda, gold, tag):
  """_resize_predictions (:530-571) run by the reference vs wlseg_resize_nearest / wlseg_resize_probs."""
  from wlseg import ops
  oh, ow = (int(x) for x in gold[f'{tag}/size'])
  got = ops.resize_decisions(torch.from_numpy(gold[f'{tag}/in_decisions']).to(cuda), oh, ow)
  assert np.array_equal(got.cpu().numpy(), gold[f'{tag}/out_decisions'])
  for k in ('l1_probabilities', 'l2_vehicle_probabilities', 'l2_human_probabilities'):
    got = ops.resize_probabilities(torch.from_numpy(gold[f'{tag}/in_{k}']).to(cuda), oh, ow)
    np.testing.assert_allclose(got.cpu().numpy(), gold[f'{tag}/out_{k}'], rtol=0, atol=1e-6)


def test_batch_mean_iou_on_device_equals_the_reference_run(cuda, gold):
  """define_metrics.mean_iou run by the reference vs confmat kernel + estimator.mean_iou_from_cm on device tensors."""
  from wlseg import estimator as west
  from wlseg import ops
  cm = torch.zeros(20, 20, dtype=torch.int64, device=cuda)
  ops.confmat_accumulate(torch.from_numpy(gold['mean_iou/labels']).to(cuda), torch.from_numpy(gold['mean_iou/decisions']).to(cuda), 20, cm)
  got = west.mean_iou_from_cm(cm, 20)
  got = float(got.cpu()) if isinstance(got, torch.Tensor) else float(got)
  assert abs(got - float(gold['mean_iou/out'])) <= 1e-6


@pytest.mark.parametrize('name,nesterov', [('plain', False), ('nesterov', True)])
def test_sgdm_kernel_equals_the_reference_optimizer(cuda, gold, name, nesterov):
  """Three updates through the optimizer object `define_optimizer` returns (define_optimizer.py:17-20) vs
  wlseg_sgdm_step (no weight decay: the L2 term is part of the gradient in the reference)."""
  from wlseg import ops
  w = torch.from_numpy(gold['sgdm/w0']).clone().to(cuda)
  acc = torch.zeros_like(w)
  wb = torch.zeros(w.numel(), dtype=torch.bfloat16, device=cuda)
  lr = torch.full((1,), 0.01, device=cuda)
  for g in torch.from_numpy(gold['sgdm/grads']):
    ops.sgdm_step(w, g.clone().to(cuda), acc, wb, 0, lr, 0.9, nesterov, 0.0)
  np.testing.assert_allclose(w.cpu().numpy(), gold[f'sgdm/{name}'], rtol=0, atol=1e-6)


@pytest.mark.parametrize('tag', ['crop_ids', 'crop_dense', 'crop_ids_tall', 'resize_plain'])
def test_resize_crop_kernel_equals_the_reference_run(cuda, gold, tag):
  """input_pipelines/utils.py:181-247 `resize_images_and_labels` run by the reference (crop offsets recorded) vs
  wlseg_resize_crop: labels (nearest) bit-exact, images (bilinear) 1e-6."""
  from wlseg import preprocess as wpre
  img, lab = torch.from_numpy(gold[f'{tag}/images']).to(cuda), torch.from_numpy(gold[f'{tag}/labels']).to(cuda)
  target = tuple(int(x) for x in gold[f'{tag}/target'])
  preserve = bool(gold[f'{tag}/preserve'])
  off = tuple(int(x) for x in gold[f'{tag}/offset'])
  pi, pl = wpre.resize_images_and_labels(img, lab, target, preserve, offset=off)
  np.testing.assert_allclose(pi.cpu().numpy(), gold[f'{tag}/out_images'], rtol=0, atol=1e-6)
  assert np.array_equal(pl.cpu().numpy(), gold[f'{tag}/out_labels'])
  if preserve:   # drawn offsets stay inside the slack and the call is deterministic under a seeded generator
    g = torch.Generator().manual_seed(1)
    a = wpre.resize_images_and_labels(img, lab, target, True, generator=g)
    g = torch.Generator().manual_seed(1)
    b = wpre.resize_images_and_labels(img, lab, target, True, generator=g)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and tuple(a[0].shape[1:3]) == target


# ------------------------------------------------------------------------------------------------ the network
# tests/golden/reference_model_run.npz: the reference's own model() executed over tests/golden/tf_shim (+ _slim.py)
MODEL_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_model_run.npz')


@pytest.fixture(scope='module')
def model_gold():
  return np.load(MODEL_GOLD)


def _model_case(tag):
  import importlib.util
  spec = importlib.util.spec_from_file_location('make_reference_model_fixtures', os.path.join(
      os.path.dirname(os.path.abspath(__file__)), 'golden', 'make_reference_model_fixtures.py'))
  gen = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(gen)          # only main() touches /root/reference
  return gen, gen.CASES[tag], gen.case_params(tag)


def _build(cuda, tag, dtype, train=False):
  from wlseg import network
  gen, (dataset, N, H, W, _, accumulate, init_kw, flags), tfp = _model_case(tag)
  hier = _hier(dataset)
  params = network.Params(hier, cuda, psp=init_kw.get('psp', False), fov=init_kw.get('fov'),
                          upsampling=init_kw.get('upsampling', 'bilinear'), norm=init_kw.get('norm', 'batch'))
  params.load_tf_dict(tfp)
  cls = network.TrainNetwork if (train or init_kw.get('norm') == 'group') else network.Network
  return gen, hier, tfp, params, cls(params, dtype=dtype)


@pytest.mark.parametrize('tag', ['cs_eval', 'vistas_eval', 'cs_psp_fov_hybrid', 'cs_group'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_network_equals_the_reference_model_run(cuda, model_gold, tag, dtype):
  """The CUDA forward pass (wlseg.network: convolutions, folded batch norm / group norm, pooling, pyramid module,
  transposed-convolution upsampler, head) against the predictions the REFERENCE's model() produced for the same
  variables and images.  fp32 check mode: full-resolution logits 1e-4 of their maximum (north star), decisions equal
  except where two logits tie to that level (<= 0.2 % of the pixels); bf16 product path: logits rel-L2 <= 2e-2
  (5e-2 for group norm, which normalises by the statistics of the rounded tensor itself)."""
  gen, hier, tfp, params, net = _build(cuda, tag, dtype)
  images = torch.from_numpy(model_gold[f'{tag}/images'])
  out = net.predict(images.to(cuda), want=('logits', 'decisions', 'l1_decisions', 'l2_vehicle_decisions', 'l2_human_decisions'))
  torch.cuda.synchronize()
  s = gen.LOGIT_STRIDE
  worst_max, num, den = 0.0, 0.0, 0.0
  for k in ('l1_logits', 'l2_vehicle_logits', 'l2_human_logits'):
    want = torch.from_numpy(model_gold[f'{tag}/{k}'])
    got = out[k].cpu()[:, ::s, ::s]
    assert got.shape == want.shape
    worst_max = max(worst_max, float((got - want).abs().max()) / float(want.abs().max()))
    num += float((got - want).double().pow(2).sum())
    den += float(want.double().pow(2).sum())
  rel_l2 = (num / den) ** 0.5
  mism = max(float((out[k].cpu() != torch.from_numpy(model_gold[f'{tag}/{k}'].astype(np.int32))).float().mean())
             for k in ('decisions', 'l1_decisions', 'l2_vehicle_decisions', 'l2_human_decisions'))
  print(f'{tag} {dtype}: logits max-rel {worst_max:.2e}, rel-L2 {rel_l2:.2e}, decision mismatch {100 * mism:.3f} %')
  if dtype == torch.float32:
    assert worst_max <= 1e-4 and mism <= 2e-3
  else:
    assert rel_l2 <= (5e-2 if tag == 'cs_group' else 2e-2) and mism <= 0.08


def test_training_mode_network_equals_the_reference_model_run(cuda, model_gold):
  """Training-mode batch norm (batch statistics in every layer, `batch_norm_accumulate_statistics`): the fp32 check mode's
  forward_train against the reference run - logits 1e-3 of their maximum (70 positions per channel in the deepest
  layers: last-bit differences are amplified layer after layer, the oracle itself is at 1e-4 here), and the moving
  statistics it leaves behind against moving - (1 - decay) * (moving - statistic) with the reference run's batch mean
  and Bessel-corrected variance."""
  tag = 'cs_train_bn'
  gen, hier, tfp, params, net = _build(cuda, tag, torch.float32, train=True)
  images = torch.from_numpy(model_gold[f'{tag}/images'])
  N, H, W, _ = images.shape
  low = net.forward_train(images.to(cuda))
  out = net.head(low, H, W, ('logits', 'decisions'))
  torch.cuda.synchronize()
  s = gen.LOGIT_STRIDE
  for k in ('l1_logits', 'l2_vehicle_logits', 'l2_human_logits'):
    want = torch.from_numpy(model_gold[f'{tag}/{k}'])
    got = out[k].cpu()[:, ::s, ::s]
    err = float((got - want).abs().max()) / float(want.abs().max())
    print(f'{k}: {err:.2e}')
    assert err <= 1e-3, k
  assert float((out['decisions'].cpu() != torch.from_numpy(model_gold[f'{tag}/decisions'].astype(np.int32))).float().mean()) <= 5e-3
  back = params.to_tf_dict()
  scopes = sorted({k.split('/update/')[1].rsplit('/', 1)[0] for k in model_gold.files if k.startswith(f'{tag}/update/')})
  assert len(scopes) >= 3
  for sc in scopes:
    mean = torch.from_numpy(model_gold[f'{tag}/update/{sc}/mean'])
    var = torch.from_numpy(model_gold[f'{tag}/update/{sc}/unbiased_variance'])
    want_mean = tfp[f'{sc}/moving_mean'] - 0.1 * (tfp[f'{sc}/moving_mean'] - mean)
    want_var = tfp[f'{sc}/moving_variance'] - 0.1 * (tfp[f'{sc}/moving_variance'] - var)
    em = float((back[f'{sc}/moving_mean'].cpu() - want_mean).abs().max()) / (float(want_mean.abs().max()) + 1e-6)
    ev = float((back[f'{sc}/moving_variance'].cpu() - want_var).abs().max()) / float(want_var.abs().max())
    assert em <= 1e-3 and ev <= 1e-3, (sc, em, ev)


# ------------------------------------------------------------------------------------------------ the TRAIN branch
@pytest.fixture(scope='module')
def train_gold():
  from tests import test_reference_fixtures as cpu_side
  return np.load(cpu_side.TRAIN_GOLD)


@pytest.mark.parametrize('tag,dtype', [('cs_mixed_sgdm_ema', 'fp32'), ('cs_mixed_sgdm_ema', 'bf16'),
                                       ('cs_strong_nesterov_poly', 'fp32'), ('cs_strong_nesterov_poly', 'bf16'),
                                       ('vistas_mixed_sgdm', 'fp32'), ('vistas_mixed_sgdm', 'bf16'),
                                       ('cs_psp_fov_hybrid', 'fp32'), ('cs_group_norm', 'fp32')])
def test_trainer_equals_the_reference_training_run(cuda, train_gold, tag, dtype):
  """The reference's `define_estimator` TRAIN branch (define_estimator_hierarchical.py:77-159: model() in training mode,
  define_losses, EMA in UPDATE_OPS, define_optimizer, create_train_op), executed by the reference itself for 3 / 2
  optimizer steps (tests/golden/make_reference_train_fixtures.py), against the product's `Trainer.step` on the same
  initial variables and batches - no oracle in between.  Compared: every step's total / l1 / l2_vehicle / l2_human /
  regularisation loss, then the state the session is left with under its TF names (`checkpoints.export_train_state`):
  variables, moving statistics, Momentum slots, EMA shadows.
  fp32 check mode: first-step losses 1e-4, later steps 5e-4 (l2 heads 5e-3), update cosine >= 0.999, norms within 2 %.
  Measured: losses within 1e-6 relative at every step, update cosine 0.9998 / 1.0000, norms within 1.2e-3.
  bf16 product path (tcgen05 convolutions, bf16 storage) against the SAME fp32 run: losses 2e-2 (north star; measured
  5e-3), l2 heads of later steps 1e-1, update cosine >= 0.90 (measured 0.946 on conv1/weights, the tensor with the
  whole network's storage roundings behind its gradient), norms within 15 % (measured 2-7 %), moving statistics 2e-2."""
  from tests import test_reference_fixtures as cpu_side
  from wlseg import checkpoints, network, trainer as wtrainer
  gen, (dataset, n_pp, n_pb, n_pi, H, W, steps, opt), batches = cpu_side.train_case_batches(train_gold, tag)
  initial = gen.case_params(tag)
  hier = _hier(dataset)
  params = network.Params(hier, cuda, **opt.get('model', ({}, {}))[0])     # psp / fov / upsampling / norm
  params.load_tf_dict(initial)
  settings = type('S', (), dict(momentum=opt['momentum'], use_nesterov=opt['use_nesterov'], optimizer=opt['optimizer'],
                                regularization_weight=opt['regularization_weight'], batch_norm_decay=opt['batch_norm_decay'],
                                distribute=False, ema_decay=opt['ema_decay']))
  tr = wtrainer.Trainer(params, settings, dtype=torch.float32 if dtype == 'fp32' else torch.bfloat16, use_graph=False)
  rows = []
  for i, (images, labels) in enumerate(batches):
    assert tr.global_step == int(train_gold[f'{tag}/step{i}/global_step_before'])
    lr = cpu_side.reference_lr(train_gold, tag, opt, tr.global_step)
    out = tr.step({'proimages': images.to(cuda)}, {k: v.to(cuda) for k, v in labels.items()}, lr).cpu()
    rows.append([float(out[0]), float(out[2]), float(out[3]), float(out[4]), float(out[5])])
    assert abs(float(out[1]) + float(out[5]) - float(out[0])) <= 1e-4 * abs(float(out[0]))   # total = segmentation + reg
  torch.cuda.synchronize()
  assert tr.global_step == int(train_gold[f'{tag}/global_step'])
  state = checkpoints.export_train_state(params, tr)
  variables = {k: v for k, v in state.items() if k in initial}
  momentum = {k: state[checkpoints.momentum_name(k)] for k in initial if checkpoints.momentum_name(k) in state}
  ema = {k: state[checkpoints.ema_name(k)] for k in initial if checkpoints.ema_name(k) in state}
  tol = dict(first_tol=1e-4, later_tol=5e-4, cos_min=0.999, norm_tol=1e-2, moving_tol=2e-3) if dtype == 'fp32' else \
      dict(first_tol=2e-2, later_tol=2e-2, cos_min=0.90, norm_tol=1.5e-1, moving_tol=2e-2)
  cpu_side.compare_train_state(train_gold, tag, gen, opt, initial, variables, momentum, ema, rows, **tol)


# ------------------------------------------------------------------------------------------------ EVAL / PREDICT branches
@pytest.fixture(scope='module')
def eval_gold():
  from tests import test_reference_fixtures as cpu_side
  return np.load(cpu_side.EVAL_GOLD)


@pytest.mark.parametrize('tag', ['eval_cs_same_size', 'eval_cs_labels_2x', 'eval_vistas_labels_odd'])
@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_system_evaluate_equals_the_reference_eval_run(cuda, eval_gold, tmp_path, tag, dtype):
  """The reference's EVAL branch (define_estimator_hierarchical.py:160-201) run by the reference itself over two
  batches, against the product's WHOLE evaluation stack on the same batches: evaluate.py's settings ->
  SemanticSegmentation.evaluate() (cid map from the problem definition, `_replacevoids`, void row / column trimmed,
  system_factory.py:400-405) -> Estimator.evaluate -> wlseg_head_confmat / resize + wlseg_confmat_accumulate, with the
  weights restored from a checkpoint file under their TF names (predict_saver).
  fp32 check mode: the streaming confusion matrix is the reference's, entry by entry.  bf16 product path: the matrices
  may differ where an arg-max is a near-tie - at most 2 % of the pixels moved (north star: decisions >= 98 %)."""
  from tests import test_reference_fixtures as cpu_side
  from wlseg import checkpoints, problem_defs, settings as wsettings
  from wlseg.system_factory import SemanticSegmentation
  gen = cpu_side._eval_gen()
  dataset, nbatches, N, H, W, LH, LW = gen.EVAL_CASES[tag]
  ckpt = checkpoints.save_file(os.path.join(str(tmp_path), 'model.ckpt-7.pt'), gen.case_params(dataset), 7)
  argv = [str(tmp_path), str(N * nbatches), problem_defs.default_path(dataset), 'unused', dataset, '--Nb', str(N),
          '--height_feature_extractor', str(H), '--width_feature_extractor', str(W), '--dtype', dtype, '--ckpt_path', ckpt]
  st = wsettings.eval_extra_args(wsettings.build_parser(wsettings.EVAL).parse_args(argv))
  st.device, st.rank, st.world_size = 'cuda:0', 0, 1

  def input_fn(config, params):
    for b in range(nbatches):
      yield ({'proimages': torch.from_numpy(eval_gold[f'{tag}/batch{b}/images'])},
             {'prolabels': torch.from_numpy(eval_gold[f'{tag}/batch{b}/prolabels'].astype(np.int32))})

  system = SemanticSegmentation({'eval': input_fn}, None, st)
  assert system.settings.training_cids2evaluation_cids == eval_gold[f'{tag}/training_cids2evaluation_cids'].tolist()
  metrics = system.evaluate()
  assert len(metrics) == 1 and metrics[0]['global_step'] == 7 and metrics[0]['steps'] == nbatches and metrics[0]['loss'] == 0.0
  ref = eval_gold[f'{tag}/confusion_matrix'].astype(np.int64)
  got = metrics[0]['confusion_matrix']
  assert got.dtype == np.int32 and got.shape == (ref.shape[0] - 1, ref.shape[1] - 1)      # void trimmed
  full = metrics[0]['confusion_matrix_int64']
  assert int(full.sum()) == int(ref.sum()) == N * nbatches * LH * LW
  assert np.array_equal(full.sum(1), ref.sum(1))                                          # label histogram: exact always
  moved = int(np.abs(full - ref).sum()) // 2
  print(f'{tag} {dtype}: {moved} of {int(ref.sum())} pixels in another cell')
  if dtype == 'fp32':
    assert np.array_equal(full, ref) and np.array_equal(got, ref[:-1, :-1])
  else:
    assert moved <= 0.02 * ref.sum(), moved


@pytest.mark.parametrize('tag', ['predict_cs_system_size', 'predict_cs_raw_size'])
def test_estimator_predict_equals_the_reference_predict_run(cuda, eval_gold, tag):
  """The reference's PREDICT branch (:204-237) run by the reference itself, against Estimator.predict (fp32 check mode):
  output size (height_system x width_system, or the raw image's), decisions (nearest), probabilities (bilinear) 1e-4,
  raw images / paths passed through."""
  import argparse
  from tests import test_reference_fixtures as cpu_side
  from wlseg import estimator as est, network
  gen = cpu_side._eval_gen()
  dataset, N, H, W, system, raw = gen.PREDICT_CASES[tag]
  s = argparse.Namespace(dtype='fp32', stride_feature_extractor=8, psp_module=False, height_system=system[0],
                         width_system=system[1], replace_voids=False, batch_norm_decay=1.0)
  e = est.Estimator(s, _hier(dataset), device=cuda)
  e.params.load_tf_dict(gen.case_params(dataset))
  e.net = network.Network(e.params, dtype=torch.float32)
  features = {'proimages': torch.from_numpy(eval_gold[f'{tag}/images'])}
  keys = ['decisions', *gen.PROB_KEYS]
  if raw is not None:
    features['rawimages'] = torch.zeros(N, raw[0], raw[1], 3, dtype=torch.uint8)
    features['rawimagespaths'] = ['a.png'] * N
    keys += ['rawimages', 'rawimagespaths']
  outs = list(e.predict([(features, None)], keys))
  assert len(outs) == N and sorted(outs[0].keys()) == str(eval_gold[f'{tag}/prediction_keys']).split('\n')
  oh, ow = (int(v) for v in eval_gold[f'{tag}/size'])
  for i, ex in enumerate(outs):
    assert ex['decisions'].shape == (oh, ow)
    dis = float((ex['decisions'] != eval_gold[f'{tag}/decisions'][i]).mean())
    assert dis <= 1e-3, dis       # fp32 on both sides: near-ties only (measured 0)
    for k in gen.PROB_KEYS:
      ref = eval_gold[f'{tag}/{k}'][i]
      assert np.abs(ex[k][::gen.PROB_STRIDE, ::gen.PROB_STRIDE] - ref).max() <= 1e-4, k
