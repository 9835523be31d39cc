"""The oracle (and the product's host-side helpers) against vectors produced by RUNNING THE REFERENCE'S OWN PYTHON.

tests/golden/reference_run.npz is written by tests/golden/make_reference_fixtures.py, which imports
/root/reference/code over a small TensorFlow-1.12 API emulation (tests/golden/tf_shim) and calls the reference's own
`define_losses`, `_segment_sum`, `_generate_rla`, `_map_predictions_to_new_cids`, `_resize_predictions`,
`_replace_voids`, `mean_iou`, `define_optimizer`, `_replacevoids`, `get_temp_Nb`,
`print_metrics_from_confusion_matrix`.  This pins the oracle's restatement of the reference's Python (tables,
gathers, masks, weights, normalisation, remaps, resize calls); TensorFlow's own kernels stay restated.
"""

import importlib
import io
import os

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import metrics as ometrics
from oracle import optimizer as oopt
from oracle import tfops
from oracle import weak_labels as oweak

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_run.npz')
LOSS_CASES = ['losses_cityscapes_mixed', 'losses_cityscapes_strong', 'losses_vistas_mixed', 'losses_vistas_strong']


@pytest.fixture(scope='module')
def gold():
  return np.load(GOLD)


def _loss_inputs(gold, tag):
  n_pp, n_pb, n_pi, h, w = (int(x) for x in gold[f'{tag}/counts'])
  H, W = 8 * h, 8 * w
  low = [torch.from_numpy(gold[f'{tag}/lowres_{k}_logits']).clone().requires_grad_(True) for k in ('l1', 'l2_vehicle', 'l2_human')]
  labels = {'prolabels_per_pixel': torch.from_numpy(gold[f'{tag}/prolabels_per_pixel'])}
  if n_pb:
    labels['prolabels_per_bbox'] = torch.from_numpy(gold[f'{tag}/prolabels_per_bbox'])
  if n_pi:
    vec = torch.from_numpy(gold[f'{tag}/prolabels_per_image_vectors'])
    labels['prolabels_per_image'] = vec[:, None, None, :].expand(n_pi, H, W, 15).contiguous()
  return str(gold[f'{tag}/dataset']), (n_pp, n_pb, n_pi, H, W), low, labels


@pytest.mark.parametrize('tag', LOSS_CASES)
def test_oracle_losses_and_gradients_equal_the_reference_run(gold, tag):
  """define_losses_hierarchical.py:14-224 executed by the reference itself vs oracle/losses.py: the five losses and
  the gradient of `total` w.r.t. the low-resolution logits (fp32, 1e-5 / 1e-4 of the gradient's max)."""
  from oracle import network as onet
  dataset, (n_pp, n_pb, n_pi, H, W), low, labels = _loss_inputs(gold, tag)
  full = [tfops.resize_bilinear(z, H, W, align_corners=True) for z in low]
  pred = onet.compose_predictions(*full, dataset)
  assert np.array_equal(pred['l1_decisions'].numpy(), gold[f'{tag}/l1_decisions'])
  got = olosses.define_losses(pred, labels, dataset)
  for k in ('l1_segmentation', 'l2_vehicle_segmentation', 'l2_human_segmentation'):
    ref = float(gold[f'{tag}/loss_{k}'])
    assert abs(float(got[k].detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (k, float(got[k].detach()), ref)
  reg = float(gold[f'{tag}/loss_regularization'])
  assert abs(float(got['segmentation'].detach()) + reg - float(gold[f'{tag}/loss_total'])) <= 1e-5
  assert float(gold[f'{tag}/loss_l1_segmentation_hot']) == 0.0
  got['segmentation'].backward()
  for z, k in zip(low, ('l1', 'l2_vehicle', 'l2_human')):
    ref = torch.from_numpy(gold[f'{tag}/grad_lowres_{k}_logits'])
    assert float((z.grad - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-9, k


def test_rasteriser_equals_generate_rla(gold):
  """input_subset_bboxes_v2.py:74-98 `_generate_rla` run by the reference vs oracle/weak_labels.py (bit-exact), on
  the docstring's normalisation cases and on every box list of the loss fixtures."""
  h, w = (int(x) for x in gold['rla/size'])
  boxes = [(int(c),) + tuple(float(v) for v in xy) for c, xy in zip(gold['rla/cids'], gold['rla/coords']) if c >= 0]
  got = oweak.bbox_labels(boxes, h, w)
  assert np.array_equal(got, gold['rla/out'])
  assert np.all(np.abs(got.sum(-1) - 1.0) < 1e-3)          # input_subset_bboxes_v2_test.py:40-43
  for tag in LOSS_CASES:
    n_pp, n_pb, n_pi, hh, ww = (int(x) for x in gold[f'{tag}/counts'])
    for i in range(n_pb):
      boxes = [(int(c),) + tuple(float(v) for v in xy)
               for c, xy in zip(gold[f'{tag}/bbox{i}_cids'], gold[f'{tag}/bbox{i}_coords']) if c >= 0]
      assert np.array_equal(oweak.bbox_labels(boxes, 8 * hh, 8 * ww), gold[f'{tag}/prolabels_per_bbox'][i]), (tag, i)


def test_image_level_vectors_equal_generate_rla(gold):
  """input_subset_image_labels.py:73-94 `_generate_rla` run by the reference (one class, several, a duplicate + an
  unknown mid, none -> void) vs oracle/weak_labels.py::image_labels, bit-exact; the tiling is :102."""
  for cids, want in zip(gold['rla_image/cids'], gold['rla_image/out']):
    present = sorted({int(c) for c in cids if c >= 0})
    got = oweak.image_labels(present, 4, 6)
    assert got.shape == (4, 6, 15) and np.array_equal(got[0, 0], want) and np.array_equal(got, np.broadcast_to(want, got.shape))


def test_segment_sum_worked_example(gold):
  """:112-113, 219-224: half a vehicle + half a human on one pixel -> 1/2 car + 1/2 void for the vehicle head."""
  lab = torch.from_numpy(gold['segment_sum/labels'])
  got = tfops.unsorted_segment_sum_last(lab, [int(x) for x in gold['segment_sum/ids']], 7)
  assert np.array_equal(got.numpy(), gold['segment_sum/out'])
  assert got[0, 0, 0].tolist() == [0.5, 0.0, 0.0, 0.0, 0.0, 0.0, 0.5]   # car -> vehicle class 0, human -> void


def test_cid_remap_equals_reference(gold):
  m = [int(x) for x in gold['remap/map']]
  assert np.array_equal(ometrics.map_decisions_to_new_cids(gold['remap/decisions'], m), gold['remap/out_decisions'])
  np.testing.assert_allclose(ometrics.map_probabilities_to_new_cids(gold['remap/probs'], m), gold['remap/out_probs'], rtol=0, atol=1e-7)
  m = [int(x) for x in gold['remap_cs/map']]
  assert np.array_equal(ometrics.map_decisions_to_new_cids(gold['remap_cs/decisions'], m), gold['remap_cs/out_decisions'])
  assert bool(gold['remap_cs/probs_untouched'])   # 14 channels vs a 20-entry map: the reference's try/except skips them
  from wlseg import estimator as west
  assert west._replacevoids([int(x) for x in gold['replacevoids/in']]) == [int(x) for x in gold['replacevoids/out']]
  assert ometrics.replacevoids([int(x) for x in gold['replacevoids/in']]) == [int(x) for x in gold['replacevoids/out']]


@pytest.mark.parametrize('tag', ['resize_up', 'resize_down'])
def test_resize_predictions_equals_reference(gold, tag):
  oh, ow = (int(x) for x in gold[f'{tag}/size'])
  got = tfops.resize_nearest(torch.from_numpy(gold[f'{tag}/in_decisions'])[..., None], oh, ow, align_corners=True)[..., 0]
  assert np.array_equal(got.numpy(), gold[f'{tag}/out_decisions'])
  for k in ('l1_probabilities', 'l2_vehicle_probabilities', 'l2_human_probabilities'):
    got = tfops.resize_bilinear(torch.from_numpy(gold[f'{tag}/in_{k}']), oh, ow, align_corners=True)
    np.testing.assert_allclose(got.numpy(), gold[f'{tag}/out_{k}'], rtol=0, atol=1e-6)


def test_replace_voids_flat_classifier(gold):
  """:573-630 on the key set it accepts: a void decision (last channel) becomes the runner-up."""
  p, d = gold['replace_voids/probs'], gold['replace_voids/decisions']
  best_non_void = np.argmax(p[..., :-1], -1)
  want = np.where(d == p.shape[-1] - 1, best_non_void, d).astype(np.int32)
  assert np.array_equal(want, gold['replace_voids/out_decisions'])


def test_batch_mean_iou_equals_reference(gold):
  got = ometrics.batch_mean_iou(gold['mean_iou/labels'], gold['mean_iou/decisions'], 20)
  assert abs(float(got) - float(gold['mean_iou/out'])) <= 1e-6
  from wlseg import estimator as west
  cm = torch.from_numpy(ometrics.confusion_matrix(gold['mean_iou/labels'], gold['mean_iou/decisions'], 20))
  assert abs(float(west.mean_iou_from_cm(cm, 20)) - float(gold['mean_iou/out'])) <= 1e-6


def test_schedules_and_momentum_equal_reference(gold):
  steps = [int(s) for s in gold['lr/steps']]
  b, v = [8 * 743, 15 * 743], [0.01, 0.005, 0.0025]
  assert [oopt.piecewise_constant(s, b, v) for s in steps] == gold['lr/piecewise'].tolist()
  got = [oopt.polynomial_decay(0.01, s, 17 * 743, 0.0001, 0.9) for s in steps]
  np.testing.assert_allclose(got, gold['lr/polynomial'], rtol=1e-12)
  # the product's host-side schedule (wlseg/estimator.py learning_rate)
  from wlseg import estimator as west

  class P:
    learning_rate_schedule, learning_rate_boundaries, learning_rate_values = 'piecewise_constant', b, v
  assert [west.learning_rate(P, s) for s in steps] == gold['lr/piecewise'].tolist()

  class Q:
    learning_rate_schedule, learning_rate_initial, learning_rate_final = 'polynomial_decay', 0.01, 0.0001
    learning_rate_power, num_training_steps = 0.9, 17 * 743
  np.testing.assert_allclose([west.learning_rate(Q, s) for s in steps], gold['lr/polynomial'], rtol=1e-12)
  for name, nesterov in (('plain', False), ('nesterov', True)):
    w = torch.from_numpy(gold['sgdm/w0']).clone()
    acc = torch.zeros_like(w)
    for g in torch.from_numpy(gold['sgdm/grads']):
      w, acc = oopt.momentum_step(w, g, acc, 0.01, 0.9, nesterov)
    np.testing.assert_allclose(w.numpy(), gold[f'sgdm/{name}'], rtol=0, atol=1e-7)


def test_host_helpers_equal_reference(gold):
  from wlseg import estimator as west
  from wlseg import metrics as wmetrics

  class P:
    distribute = False
  assert west.get_temp_Nb(P, 8) == int(gold['temp_nb/out'][0])
  assert gold['m1_1/out'].tolist() == [-1.0, -0.5, 0.0, 1.0]
  cm = gold['print_metrics/cm']
  buf = io.StringIO()
  with np.errstate(all='ignore'):
    wmetrics.print_metrics_from_confusion_matrix(cm, printfile=buf, summary=True)
  assert buf.getvalue() == str(gold['print_metrics/summary'])
  m = ometrics.metrics_from_confusion_matrix(cm.astype(np.int64))
  assert f"Mean iou (ignoring accuracies' nans but including ious' 0s): {m['mean_iou']:5.2f}" in str(gold['print_metrics/summary'])


CROP_CASES = ['crop_ids', 'crop_dense', 'crop_ids_tall', 'resize_plain']


@pytest.mark.parametrize('tag', CROP_CASES)
def test_resize_and_crop_equals_reference(gold, tag):
  """input_pipelines/utils.py:181-247 run by the reference (random crop offsets recorded) vs oracle/preprocess.py and
  the product's size arithmetic (wlseg/preprocess.py resized_size)."""
  from oracle import preprocess as opre
  from wlseg import preprocess as wpre
  img, lab = torch.from_numpy(gold[f'{tag}/images']), torch.from_numpy(gold[f'{tag}/labels'])
  target = tuple(int(x) for x in gold[f'{tag}/target'])
  preserve = bool(gold[f'{tag}/preserve'])
  off = tuple(int(x) for x in gold[f'{tag}/offset'])
  pi, pl = opre.resize_images_and_labels(img, lab, target, preserve, off)
  np.testing.assert_allclose(pi.numpy(), gold[f'{tag}/out_images'], rtol=0, atol=1e-6)
  assert np.array_equal(pl.numpy(), gold[f'{tag}/out_labels'])
  RH, RW = wpre.resized_size(img.shape[1], img.shape[2], target, preserve)
  assert 0 <= off[0] <= RH - target[0] and 0 <= off[1] <= RW - target[1]
  if preserve:   # tight fit: one of the two axes has no slack
    assert RH == target[0] or RW == target[1]


# ------------------------------------------------------------------------------------------------ the network
# tests/golden/reference_model_run.npz: the reference's own model() executed over tests/golden/tf_shim (+ _slim.py)
MODEL_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_model_run.npz')
MODEL_CASES = ['cs_eval', 'cs_train_bn', 'vistas_eval', 'cs_psp_fov_hybrid', 'cs_group', 'vistas_odd_size', 'cs_odd_size_train_bn']


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


@pytest.mark.parametrize('tag', MODEL_CASES)
def test_oracle_network_equals_the_reference_model_run(model_gold, tag):
  """oracle/network.py::Net against the predictions the REFERENCE's model() returned (feature_extractor, arg scope,
  adaptation bottlenecks, logits + normaliser, upsampler, pyramid / field-of-view / hybrid-upsampling / group-norm
  variants, softmax / argmax / decision composition): logits 1e-5 of their maximum on the stored grid, every
  decision map equal."""
  from oracle import network as onet
  gen, (dataset, N, H, W, train, accumulate, init_kw, flags), tfp = _model_case(tag)
  assert abs(gen.checksum(tfp) - float(model_gold[f'{tag}/params_checksum'])) <= 1e-6 * float(model_gold[f'{tag}/params_checksum'])
  images = torch.from_numpy(model_gold[f'{tag}/images'])
  assert torch.equal(images, gen.case_images(tag))
  net = onet.Net(tfp, dataset, training=accumulate, psp=init_kw.get('psp', False), fov=init_kw.get('fov'),
                 upsampling=init_kw.get('upsampling', 'bilinear'), norm=init_kw.get('norm', 'batch'))
  with torch.no_grad():
    pred = net.forward(images)
  s = gen.LOGIT_STRIDE
  # inference-mode normalisation: 1e-5 and identical decisions.  Normalising by the statistics of the tensor itself
  # (training-mode batch norm over 2 x 5 x 7 positions, group norm per sample) amplifies the last-bit differences of
  # two fp32 evaluation orders layer after layer: 1e-3, and decisions may differ where two logits tie to that level
  exact = tag not in ('cs_train_bn', 'cs_group', 'cs_odd_size_train_bn')
  tol = 1e-5 if exact else 1e-3
  worst = 0.0
  for k in ('l1_logits', 'l2_vehicle_logits', 'l2_human_logits'):
    want = torch.from_numpy(model_gold[f'{tag}/{k}'])
    got = pred[k][:, ::s, ::s]
    assert got.shape == want.shape
    err = float((got - want).abs().max()) / float(want.abs().max())
    worst = max(worst, err)
    assert err <= tol, (k, err)
  for k in ('decisions', 'l1_decisions', 'l2_vehicle_decisions', 'l2_human_decisions'):
    want = torch.from_numpy(model_gold[f'{tag}/{k}'].astype(np.int32))
    got = pred[k].to(torch.int32)
    if exact:
      assert torch.equal(got, want), k
    else:
      assert float((got != want).float().mean()) <= 5e-3, k
  print(f'{tag}: worst logits error {worst:.2e} of the maximum')


@pytest.mark.parametrize('tag', ['cs_eval', 'cs_psp_fov_hybrid', 'cs_group'])
def test_product_variable_names_equal_what_the_reference_model_asks_for(model_gold, tag):
  """The checkpoint variable names and shapes of the product (wlseg/arch.py, wlseg/checkpoints.py) are EXACTLY the
  variables the reference's model() requested from the variable store while it ran (slim's scoping rules restated in
  tests/golden/tf_shim/tensorflow/_slim.py): e.g. adaptation_module/l1_features/conv1/weights - an explicit `scope`
  replaces resnet_v1.bottleneck's default 'bottleneck_v1' - and the default scopes Conv .. Conv_4 / Conv2d_transpose .. _2."""
  import types
  from wlseg import arch, checkpoints as ck
  gen, (dataset, N, H, W, train, accumulate, init_kw, flags), tfp = _model_case(tag)
  up = init_kw.get('upsampling', 'bilinear')
  specs = arch.conv_specs((14, 7, 3), psp=init_kw.get('psp', False), fov=init_kw.get('fov'), upsampling=up)
  p = types.SimpleNamespace(specs=specs, norm=init_kw.get('norm', 'batch'),
                            plain=tuple(arch.UPSAMPLING_SCOPES) if up == 'hybrid' else ())
  mine = dict(ck.model_variables(p))
  asked = str(model_gold[f'{tag}/variables']).split('\n')
  assert len(set(asked)) == len(asked)
  assert set(mine) == set(asked), sorted(set(mine) ^ set(asked))[:6]
  for name, shape in mine.items():
    assert tuple(tfp[name].shape) == tuple(shape), name
  # every convolution kernel carries an L2 regulariser (module_arg_scope, _create_upsampler); its weight is
  # --regularization_weight in TRAIN mode (the evaluation graph keeps module_arg_scope's default, where it is unused)
  reg = [l.split() for l in str(model_gold[f'{tag}/regularized']).split('\n')]
  assert {n for n, _ in reg} == {n for n in mine if n.endswith('/weights')}
  reg = [l.split() for l in str(model_gold['cs_train_bn/regularized']).split('\n')]
  assert len(reg) == 66 and {float(v) for _, v in reg} == {0.00017}


def test_arg_scope_constants_of_the_reference_run(model_gold):
  """module_arg_scope as it actually reached the normaliser calls (resnet50_extended_model_hierarchical.py:278-354):
  epsilon 1e-5, scale (gamma) on, decay = --batch_norm_decay in TRAIN, is_training = batch_norm_accumulate_statistics;
  group norm: 32 groups, 1 for the logits layers (`args_context(groups=1)`, :75-77)."""
  from wlseg import network
  import inspect
  calls = [l.split() for l in str(model_gold['cs_train_bn/norm_calls']).split('\n')]
  assert len(calls) == 66 and {c[1] for c in calls} == {'batch'}
  assert {(float(c[2]), float(c[3]), c[4], c[5]) for c in calls} == {(0.9, 1e-5, 'True', 'True')}
  calls = [l.split() for l in str(model_gold['cs_eval/norm_calls']).split('\n')]
  assert {(float(c[3]), c[4], c[5]) for c in calls} == {(1e-5, 'True', 'False')}
  sig = inspect.signature(network.TrainNetwork.__init__).parameters
  assert sig['bn_decay'].default == 0.9 and sig['eps'].default == 1e-5
  calls = [l.split() for l in str(model_gold['cs_group/norm_calls']).split('\n')]
  groups = {c[0].rsplit('/', 1)[0]: int(c[5]) for c in calls}
  assert {g for s, g in groups.items() if s.startswith('softmax_classifier/')} == {1}
  assert {g for s, g in groups.items() if not s.startswith('softmax_classifier/')} == {32}
  assert {float(c[3]) for c in calls} == {1e-5}


def test_moving_statistics_of_the_reference_run(model_gold):
  """Training-mode batch norm of the reference run: the statistics a layer would push into its moving averages (batch
  mean, Bessel-corrected batch variance - [TF-1.12] fused batch norm) against the oracle's `new_moving`."""
  from oracle import network as onet
  gen, (dataset, N, H, W, train, accumulate, init_kw, flags), tfp = _model_case('cs_train_bn')
  net = onet.Net(tfp, dataset, training=True, bn_decay=0.9)
  with torch.no_grad():
    net.forward(torch.from_numpy(model_gold['cs_train_bn/images']))
  scopes = sorted({k.split('/update/')[1].rsplit('/', 1)[0] for k in model_gold.files if k.startswith('cs_train_bn/update/')})
  assert len(scopes) >= 3
  for sc in scopes:
    mean = torch.from_numpy(model_gold[f'cs_train_bn/update/{sc}/mean'])
    var = torch.from_numpy(model_gold[f'cs_train_bn/update/{sc}/unbiased_variance'])
    decay = float(model_gold[f'cs_train_bn/update/{sc}/decay'])
    assert decay == 0.9
    want_mean = tfp[f'{sc}/moving_mean'] - (1 - decay) * (tfp[f'{sc}/moving_mean'] - mean)
    want_var = tfp[f'{sc}/moving_variance'] - (1 - decay) * (tfp[f'{sc}/moving_variance'] - var)
    got_mean, got_var = net.new_moving[f'{sc}/moving_mean'], net.new_moving[f'{sc}/moving_variance']
    assert float((got_mean - want_mean).abs().max()) <= 1e-5 * float(want_mean.abs().max()) + 1e-6, sc
    assert float((got_var - want_var).abs().max()) <= 1e-5 * float(want_var.abs().max()) + 1e-6, sc


# ------------------------------------------------------------------------------------------------ the driver layer
# tests/golden/reference_driver_run.json: the reference's own argument parsers, train.py::_add_extra_args,
# evaluate.py::_add_extra_args and SemanticSegmentation.__init__ / .train() / .evaluate() executed over tests/golden/tf_shim
DRIVER_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_driver_run.json')
# values that name files of the reference's author's machine / checkout, not behaviour
_PATH_KEYS = {'init_ckpt_path', 'training_problem_def_path', 'tfrecords_path', 'tfrecords_path_per_pixel', 'tfrecords_path_per_bbox',
              'tfrecords_path_per_image', 'evaluation_problem_def_path', 'inference_problem_def_path', 'ckpt_path'}


class _FakeEstimator:
  def __init__(self, system, calls):
    self.system, self.calls = system, calls

  def train(self, batches, max_steps):
    self.calls.append(('train', max_steps))

  def evaluate(self, batches, num_classes, lut=None):
    # (the product's estimator consumes the batches the input function yields - settings.num_eval_steps of them)
    self.calls.append(('evaluate', num_classes))
    C = self.system.settings.output_Nclasses
    cm = ((np.arange(C * C, dtype=np.int64).reshape(C, C) % 7) + np.eye(C, dtype=np.int64) * 50).astype(np.int32)
    return {'global_step': 1234, 'loss': 0.5, 'confusion_matrix': cm}


def _load_driver_gold():
  import json
  with open(DRIVER_GOLD) as fp:
    return json.load(fp)


@pytest.mark.parametrize('tag', ['train_cityscapes_defaults', 'train_vistas_defaults', 'train_cityscapes_flags',
                                 'train_cityscapes_void_poly'])
def test_train_driver_settings_equal_the_reference_run(tag, tmp_path, monkeypatch):
  """wlseg.settings (CLI) + train_extra_args + wlseg.system_factory.SemanticSegmentation.__init__ / .train() on the argv
  the reference run was given: every attribute the REFERENCE left on `system.settings` (parsed flags with their defaults,
  train.py's hard overrides, derived class count, steps per epoch, total steps, learning-rate boundaries in steps and the
  values, checkpoint cadence, the EMA switch under --distribute) has the same value here, and the estimator is asked
  for the same number of steps."""
  from wlseg import settings as wsettings
  from wlseg import system_factory as sf
  gold = _load_driver_gold()[tag]
  st = wsettings.build_parser(wsettings.TRAIN).parse_args([str(tmp_path / 'log')] + gold['argv'])
  for k, v in gold['parsed'].items():
    if k not in _PATH_KEYS:
      assert getattr(st, k) == v, f'parsed flag {k}: {getattr(st, k)!r} != {v!r}'
  wsettings.train_extra_args(st)
  st.rank, st.world_size = 0, 1
  calls = []
  system = sf.SemanticSegmentation({'train': lambda config, params: iter(())}, None, st)
  monkeypatch.setattr(system, '_create_estimator',
                      lambda *a, **k: setattr(system, '_estimator', _FakeEstimator(system, calls)))
  system.train()
  mine = vars(system.settings)
  for k, v in gold['settings'].items():
    if k in _PATH_KEYS:
      continue
    assert k in mine, f'the reference sets settings.{k}, the product does not'
    assert mine[k] == v, f'settings.{k}: {mine[k]!r} != {v!r} (reference)'
  want_steps = [kw['max_steps'] for n, kw in gold['calls'] if n == 'train']
  assert calls == [('train', want_steps[0])]


@pytest.mark.parametrize('tag', ['eval_cityscapes', 'eval_vistas', 'eval_cityscapes_train_void'])
def test_evaluate_driver_settings_equal_the_reference_run(tag, tmp_path, monkeypatch):
  """The evaluation side of the same: evaluate.py's flags + overrides, derived step counts, the id map to evaluation
  classes, `eval_NN` directory naming, and the confusion matrix handed back for a given raw one (void row / column
  trimmed, system_factory.py:400-405)."""
  from wlseg import settings as wsettings
  from wlseg import system_factory as sf
  from wlseg import problem_defs
  gold = _load_driver_gold()[tag]
  argv = list(gold['argv'])
  problem_defs.write_all()
  argv[1] = problem_defs.default_path(argv[3])    # the same problem definition, at the product's location
  st = wsettings.build_parser(wsettings.EVAL).parse_args([str(tmp_path / 'log')] + argv)
  for k, v in gold['parsed'].items():
    if k not in _PATH_KEYS:
      assert getattr(st, k) == v, f'parsed flag {k}: {getattr(st, k)!r} != {v!r}'
  st = wsettings.eval_extra_args(st)
  st.rank, st.world_size = 0, 1
  st.synthetic = True
  os.makedirs(st.log_dir, exist_ok=True)
  calls = []
  system = sf.SemanticSegmentation({'eval': lambda config, params: iter(())}, None, st)
  monkeypatch.setattr(system, '_create_estimator',
                      lambda *a, **k: setattr(system, '_estimator', _FakeEstimator(system, calls)))
  import contextlib
  with contextlib.redirect_stdout(io.StringIO()):
    metrics = system.evaluate()
  mine = vars(system.settings)
  for k, v in gold['settings'].items():
    if k in _PATH_KEYS:
      continue
    assert k in mine, f'the reference sets settings.{k}, the product does not'
    assert mine[k] == v, f'settings.{k}: {mine[k]!r} != {v!r} (reference)'
  assert os.path.basename(system.settings.eval_res_dir) == gold['eval_res_dir_name']
  assert system.settings.num_eval_steps == [kw['steps'] for n, kw in gold['calls'] if n == 'evaluate'][0]
  assert calls == [('evaluate', gold['settings']['output_Nclasses'])]
  cm = np.asarray(metrics[0]['confusion_matrix'])
  assert list(cm.shape) == gold['returned_cm_shape'] and int(cm.sum()) == gold['returned_cm_sum']


# ------------------------------------------------------------------------------------------------ the TRAIN branch
TRAIN_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_train_run.npz')
TRAIN_CASES = ['cs_mixed_sgdm_ema', 'cs_strong_nesterov_poly', 'vistas_mixed_sgdm', 'cs_psp_fov_hybrid', 'cs_group_norm',
               'cs_odd_size_momentum']


@pytest.fixture(scope='module')
def train_gold():
  return np.load(TRAIN_GOLD)


def train_case_batches(train_gold, tag):
  """-> (generator module, case tuple, [(images, labels dict with DENSE weak labels)] per step)."""
  import importlib
  gen = importlib.import_module('tests.golden.make_reference_train_fixtures')
  dataset, n_pp, n_pb, n_pi, H, W, steps, opt = gen.CASES[tag]
  batches = []
  for i in range(steps):
    images = torch.from_numpy(train_gold[f'{tag}/step{i}/images'])
    labels = {'prolabels_per_pixel': torch.from_numpy(train_gold[f'{tag}/step{i}/prolabels_per_pixel'].astype(np.int32))}
    if n_pb:
      dense = []
      for j in range(n_pb):
        coords, cids = train_gold[f'{tag}/step{i}/bbox{j}_coords'], train_gold[f'{tag}/step{i}/bbox{j}_cids']
        dense.append(torch.from_numpy(oweak.bbox_labels([(int(c),) + tuple(float(v) for v in xy) for c, xy in zip(cids, coords)], H, W)))
      labels['prolabels_per_bbox'] = torch.stack(dense)
    if n_pi:
      vec = torch.from_numpy(train_gold[f'{tag}/step{i}/image_vectors'])
      labels['prolabels_per_image'] = vec[:, None, None, :].expand(n_pi, H, W, 15).contiguous()
    batches.append((images, labels))
  return gen, gen.CASES[tag], batches


def test_generator_inputs_are_reproducible(train_gold):
  """The stored images / labels are what the committed generator's seeded `case_batches` produces."""
  import importlib
  gen = importlib.import_module('tests.golden.make_reference_train_fixtures')
  for tag in TRAIN_CASES:
    for i, (images, per_pixel, boxes, vectors) in enumerate(gen.case_batches(tag)):
      assert np.array_equal(images.numpy(), train_gold[f'{tag}/step{i}/images'])
      assert np.array_equal(per_pixel.numpy(), train_gold[f'{tag}/step{i}/prolabels_per_pixel'])
      for j, (coords, cids) in enumerate(boxes):
        assert np.array_equal(coords, train_gold[f'{tag}/step{i}/bbox{j}_coords'])
        assert np.array_equal(cids, train_gold[f'{tag}/step{i}/bbox{j}_cids'])


def compare_train_state(train_gold, tag, gen, opt, initial, variables, momentum, ema, loss_rows, first_tol, later_tol,
                        cos_min, norm_tol, moving_tol=2e-3):
  """Shared by the CPU (oracle) and GPU (product) replays of a reference training run.
  loss_rows: per step [total, l1, l2_vehicle, l2_human, regularization]; variables / momentum / ema: {TF name: tensor}.
  The UPDATES (final - initial variables, Momentum slots, shadow - initial) are compared by cosine and norm ratio."""
  for i, mine in enumerate(loss_rows):
    ref = train_gold[f'{tag}/step{i}/losses']
    # later steps: the two l2 losses are averaged over pixels selected by the l1 DECISIONS of the weak images
    # (define_losses_hierarchical.py:154-185) - a handful of flipped arg-maxes moves them discontinuously: 10x the bound
    tols = [first_tol] * 5 if i == 0 else [later_tol, later_tol, 10 * later_tol, 10 * later_tol, later_tol]
    print(tag, 'step', i, 'losses', [round(float(m), 6) for m in mine], 'reference', ref.tolist())
    for m, r, tol in zip(mine, ref, tols):
      assert abs(float(m) - r) <= tol * max(1.0, abs(r)), (tag, i, list(mine), ref.tolist())
  names = str(train_gold[f'{tag}/names']).split('\n')
  ema_names = [n for n in str(train_gold[f'{tag}/ema_names']).split('\n') if n]
  # which variables carry an EMA: the model variables without 'BatchNorm/moving' (:103-106), under the scope of :97
  assert sorted(f'exponential_moving_averages/{k}/ExponentialMovingAverage' for k in ema) == ema_names
  assert sorted(variables) == names
  for n, r in zip(names, train_gold[f'{tag}/final/momentum_checksums']):
    assert (n in momentum) == (r >= 0), ('momentum slot', n)      # one slot per TRAINABLE variable
  # every variable moved by about what the reference moved it (sum |final - initial|)
  for n, r in zip(names, train_gold[f'{tag}/final/update_checksums']):
    m = float((variables[n].double() - initial[n].double()).abs().sum())
    assert abs(m - r) <= max(5.0 * norm_tol * r, 1e-9), ('update size', n, m, r)
  worst_cos, worst_norm, bad = 1.0, 0.0, []
  for store, key, base in ((variables, 'final', initial), (momentum, 'final_momentum', None), (ema, 'final_ema', initial)):
    for n in gen.KEEP:
      if n not in initial:      # group norm: <scope>/GroupNorm/{beta,gamma}, no moving statistics
        n = n.replace('/BatchNorm/', '/GroupNorm/')
      k = f'{tag}/{key}/{n}'
      if k not in train_gold.files:
        assert n not in store or key == 'final', (key, n)
        continue
      r = torch.from_numpy(train_gold[k]).double()
      m = store[n].double().cpu()
      if '/moving_' in n:
        err = float((m - r).abs().max()) / float(r.abs().max())
        if err > moving_tol:
          bad.append((key, n, 'moving statistic', err))
        continue
      if base is not None:
        r, m = r - base[n].double(), m - base[n].double()
      cos = float((r * m).sum() / (r.norm() * m.norm()))
      ratio = float(m.norm() / r.norm())
      worst_cos, worst_norm = min(worst_cos, cos), max(worst_norm, abs(ratio - 1.0))
      if not (cos >= cos_min and abs(ratio - 1.0) <= norm_tol):
        bad.append((key, n, cos, ratio))
  print(tag, f'worst update cosine {worst_cos:.6f}, worst norm deviation {worst_norm:.2e}')
  assert not bad, bad


def test_product_state_names_equal_the_reference_training_run(train_gold):
  """The names under which the product exports its training state (wlseg/checkpoints.py) against what the reference's
  TRAIN graph created: the model variables, and an ExponentialMovingAverage shadow for exactly the variables
  define_estimator_hierarchical.py:103-106 selects, under the scope of :97; one Momentum slot per trainable variable
  (none with --optimizer SGD).  Plain, Vistas, pyramid + field-of-view + hybrid upsampling, group norm."""
  import types
  from wlseg import arch, checkpoints as ck
  gen = importlib.import_module('tests.golden.make_reference_train_fixtures')
  for tag in TRAIN_CASES:
    dataset, opt = gen.CASES[tag][0], gen.CASES[tag][7]
    kw = opt.get('model', ({}, {}))[0]
    up = kw.get('upsampling', 'bilinear')
    heads = (14, 7, 3) if dataset == 'cityscapes' else (53, 12, 5)
    p = types.SimpleNamespace(specs=arch.conv_specs(heads, psp=kw.get('psp', False), fov=kw.get('fov'), upsampling=up),
                              norm=kw.get('norm', 'batch'), plain=tuple(arch.UPSAMPLING_SCOPES) if up == 'hybrid' else ())
    names = [n for n, _ in ck.model_variables(p)]
    assert sorted(names) == str(train_gold[f'{tag}/names']).split('\n'), tag
    ema_names = [n for n in str(train_gold[f'{tag}/ema_names']).split('\n') if n]
    if opt['ema_decay'] > 0:
      assert sorted(ck.ema_name(n) for n in names if ck.has_ema(n)) == ema_names, tag
    else:
      assert ema_names == []
    slots = train_gold[f'{tag}/final/momentum_checksums']
    want = [ck.trainable(n) and opt['optimizer'] == 'SGDM' for n in sorted(names)]
    assert want == [bool(c >= 0) for c in slots], tag


def reference_lr(train_gold, tag, opt, step):
  if opt['learning_rate_schedule'] == 'piecewise_constant':
    lr = oopt.piecewise_constant(step, opt['learning_rate_boundaries'], opt['learning_rate_values'])
  else:
    lr = oopt.polynomial_decay(opt['learning_rate_initial'], step, opt['num_training_steps'],
                               opt['learning_rate_final'], opt['learning_rate_power'])
  assert abs(lr - float(train_gold[f'{tag}/step{step}/learning_rate'])) <= 1e-12
  return lr


@pytest.mark.parametrize('tag', TRAIN_CASES)
def test_oracle_training_steps_equal_the_reference_run(train_gold, tag):
  """define_estimator_hierarchical.py:77-159 executed by the reference itself (model() in training mode -> define_losses
  -> EMA in UPDATE_OPS -> define_optimizer -> create_train_op) for 3 / 2 optimizer steps vs oracle/train.py on the same
  seeded batches: every step's five losses and learning rate, then the variables, Momentum slots, EMA shadows and moving
  statistics the session is left with.  fp32 on both sides over independent formulations of a train-mode BN ResNet
  (NHWC tf-shim vs the oracle's own ops; the generator's header explains the conditioning): first-step losses 1e-5
  relative, later steps 1e-4 (l2 heads 1e-3); updates: cosine >= 0.999, norms within 1 %; moving statistics 2e-3 of their maximum."""
  from oracle import train as otrain
  gen, (dataset, n_pp, n_pb, n_pi, H, W, steps, opt), batches = train_case_batches(train_gold, tag)
  initial = gen.case_params(tag)
  state = otrain.TrainState(initial, ema_decay=opt['ema_decay'], optimizer=opt['optimizer'])
  model_flags = opt.get('model', ({}, {}))[0]
  rows = []
  for i, (images, labels) in enumerate(batches):
    assert int(train_gold[f'{tag}/step{i}/global_step_before']) == state.global_step == i
    lr = reference_lr(train_gold, tag, opt, state.global_step)
    got = otrain.train_step(state, images, labels, dataset, lr, momentum=opt['momentum'], nesterov=opt['use_nesterov'],
                            regularization_weight=opt['regularization_weight'], bn_decay=opt['batch_norm_decay'], **model_flags)
    rows.append([got[k] for k in ('total', 'l1_segmentation', 'l2_vehicle_segmentation', 'l2_human_segmentation', 'regularization')])
  assert state.global_step == int(train_gold[f'{tag}/global_step'])
  if opt['ema_decay'] > 0:
    assert str(train_gold[f'{tag}/ema_notice']) == (f'Found {len(initial)} variables, saving exponential moving averages '
                                                    f'for {len(state.ema)} of them.')
  compare_train_state(train_gold, tag, gen, opt, initial, state.vars, state.momentum, state.ema, rows,
                      first_tol=1e-5, later_tol=1e-4, cos_min=0.999, norm_tol=1e-2)


# ------------------------------------------------------------------------------------------------ EVAL / PREDICT branches
EVAL_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_eval_run.npz')


@pytest.fixture(scope='module')
def eval_gold():
  return np.load(EVAL_GOLD)


def _eval_gen():
  import importlib
  return importlib.import_module('tests.golden.make_reference_eval_fixtures')


@pytest.mark.parametrize('tag', ['eval_cs_same_size', 'eval_cs_labels_2x', 'eval_vistas_labels_odd', 'eval_cs_no_upsampling'])
def test_oracle_eval_branch_equals_the_reference_run(eval_gold, tag):
  """define_estimator_hierarchical.py:160-201 run by the reference (inference-mode model(), cid map, nearest resize to
  the label size, streaming confusion matrix over the batches) vs the oracle's pieces composed in the same order:
  decisions of every batch and the accumulated confusion matrix, bit-exact (fp32 logits on both sides; an arg-max flip
  between the two formulations would show up here - there is none on these inputs)."""
  from oracle import network as onet
  gen = _eval_gen()
  dataset, nbatches, N, H, W, LH, LW = gen.EVAL_CASES[tag]
  t2e = eval_gold[f'{tag}/training_cids2evaluation_cids'].tolist()
  lut = ometrics.replacevoids(t2e)
  num_classes = max(lut) + 1
  assert num_classes == int(eval_gold[f'{tag}/num_classes'])
  # --upsampling_method no: the network's predictions stay at H/8 x W/8, the nearest resize carries them to the labels
  net = onet.Net(gen.case_params(dataset), dataset, training=False, upsampling='no' if tag.endswith('no_upsampling') else 'bilinear')
  cm = np.zeros((num_classes, num_classes), dtype=np.int64)
  for b in range(nbatches):
    with torch.no_grad():
      pred = net.forward(torch.from_numpy(eval_gold[f'{tag}/batch{b}/images']))
    decs = ometrics.map_decisions_to_new_cids(pred['decisions'].numpy(), t2e)
    decs = tfops.resize_nearest(torch.from_numpy(decs)[..., None], LH, LW, align_corners=True)[..., 0].numpy()
    assert np.array_equal(decs, eval_gold[f'{tag}/batch{b}/decisions']), tag
    cm += ometrics.confusion_matrix(eval_gold[f'{tag}/batch{b}/prolabels'], decs, num_classes)
  assert np.array_equal(cm, eval_gold[f'{tag}/confusion_matrix'])
  assert np.array_equal(ometrics.confusion_matrix_c(eval_gold[f'{tag}/batch0/prolabels'], eval_gold[f'{tag}/batch0/decisions'], num_classes),
                        ometrics.confusion_matrix(eval_gold[f'{tag}/batch0/prolabels'], eval_gold[f'{tag}/batch0/decisions'], num_classes))


@pytest.mark.parametrize('tag', ['predict_cs_system_size', 'predict_cs_raw_size'])
def test_oracle_predict_branch_equals_the_reference_run(eval_gold, tag):
  """define_estimator_hierarchical.py:204-237: the four supported keys resized to (height_system, width_system), or to
  the raw image's size when either is unset (raw images and paths are then passed through)."""
  from oracle import network as onet
  gen = _eval_gen()
  dataset, N, H, W, system, raw = gen.PREDICT_CASES[tag]
  oh, ow = (int(v) for v in eval_gold[f'{tag}/size'])
  assert (oh, ow) == (tuple(system) if raw is None else tuple(raw))
  keys = str(eval_gold[f'{tag}/prediction_keys']).split('\n')
  assert keys == sorted(['decisions', *gen.PROB_KEYS] + (['rawimages', 'rawimagespaths'] if raw is not None else []))
  net = onet.Net(gen.case_params(dataset), dataset, training=False)
  with torch.no_grad():
    pred = net.forward(torch.from_numpy(eval_gold[f'{tag}/images']))
  decs = tfops.resize_nearest(pred['decisions'][..., None], oh, ow, align_corners=True)[..., 0].numpy()
  assert np.array_equal(decs, eval_gold[f'{tag}/decisions'])
  for k in gen.PROB_KEYS:
    got = tfops.resize_bilinear(pred[k], oh, ow, align_corners=True).numpy()[:, ::gen.PROB_STRIDE, ::gen.PROB_STRIDE]
    assert np.abs(got - eval_gold[f'{tag}/{k}']).max() <= 1e-5, k


def test_product_restore_names_equal_the_reference_savers(eval_gold):
  """evaluate_saver / predict_saver (define_savers.py:38-69) as the reference built them in the EVAL and PREDICT runs:
  checkpoint key -> graph variable, with and without --restore_emas, against wlseg.checkpoints.predict_var_dict."""
  import types
  from wlseg import arch, checkpoints as ck
  p = types.SimpleNamespace(specs=arch.conv_specs((14, 7, 3)), norm='batch', plain=())
  for tag, emas in (('eval_cs_same_size', False), ('predict_cs_system_size', False), ('predict_cs_raw_size', True)):
    ref = dict(l.split() for l in str(eval_gold[f'{tag}/saver']).split('\n'))
    assert ref.pop('global_step') == 'global_step'
    assert ref == dict(ck.predict_var_dict(p, restore_emas=emas)), tag


# ------------------------------------------------------------------------------------------------ --cross_replica_norm
def test_oracle_cross_replica_batch_norm_equals_the_reference_layer_run():
  """CrossReplicaBatchNormalization._fused_batch_norm + _assign_moving_average (code/utils/
  cross_replica_batch_normalization.py:381-476) executed by the reference over a 2-tower emulation
  (tests/golden/make_reference_xreplica_fixtures.py) vs oracle/tfops.py::cross_replica_batch_norm: per-tower outputs,
  moving statistics after the update - including the reference's (n - 1) / n factor with the PER-TOWER n on a variance
  without Bessel's correction - and the gradients through the cross-tower moments (1e-5 of each tensor's maximum)."""
  gold = np.load(os.path.join(os.path.dirname(TRAIN_GOLD), 'reference_xreplica_run.npz'))
  xs = [torch.from_numpy(gold[f'tower{r}/x']).clone().requires_grad_(True) for r in range(2)]
  gamma = torch.from_numpy(gold['gamma']).clone().requires_grad_(True)
  beta = torch.from_numpy(gold['beta']).clone().requires_grad_(True)
  ys, mm, mv, mean, var = tfops.cross_replica_batch_norm(
      xs, gamma, beta, torch.from_numpy(gold['moving_mean_before']), torch.from_numpy(gold['moving_variance_before']),
      decay=float(gold['momentum']), eps=float(gold['epsilon']))
  loss = sum((y * torch.from_numpy(gold[f'tower{r}/w'])).sum() for r, y in enumerate(ys))
  grads = torch.autograd.grad(loss, xs + [gamma, beta])

  def close(a, b, what):
    b = torch.from_numpy(b)
    err = float((a.detach() - b).abs().max()) / float(b.abs().max())
    assert err <= 1e-5, (what, err)

  for r in range(2):
    close(ys[r], gold[f'tower{r}/y'], f'y{r}')
    close(grads[r], gold[f'tower{r}/dx'], f'dx{r}')
  close(grads[2], gold['dgamma'], 'dgamma')
  close(grads[3], gold['dbeta'], 'dbeta')
  close(mm, gold['tower0/moving_mean_after'], 'moving mean')
  close(mv, gold['tower0/moving_variance_after'], 'moving variance')
  # the quirk is visible at this sample size: Bessel's correction instead would differ by ~2/n = 2.9e-2 of the update
  n = xs[0].shape[0] * xs[0].shape[1] * xs[0].shape[2]
  bessel = torch.from_numpy(gold['moving_variance_before']) * 0.9 + 0.1 * var.detach() * n / (n - 1)
  assert float((bessel - torch.from_numpy(gold['tower0/moving_variance_after'])).abs().max()) > 1e-3


# ------------------------------------------------------------------------------------------------ warm start / train_saver
@pytest.mark.parametrize('tag,psp', [('warm_start', False), ('warm_start_psp', True)])
def test_product_warm_start_equals_the_reference_graph(train_gold, tag, psp):
  """The TRAIN graph built by the reference with --init_ckpt_path (define_estimator -> replace_initializers,
  define_initializers.py:72-131; train_saver, define_savers.py:3-36): tf.global_variables() of that graph - model
  variables, global_step, `exponential_moving_averages/.../ExponentialMovingAverage` shadows, `train_ops/.../Momentum`
  slots -, the checkpoint-name -> graph-variable map handed to tf.train.init_from_checkpoint, and the saved set,
  against wlseg.checkpoints.global_variables / match_init_checkpoint / export names."""
  import types
  from wlseg import arch, checkpoints as ck
  p = types.SimpleNamespace(specs=arch.conv_specs((14, 7, 3), psp=psp), norm='batch', plain=())
  ref_globals = str(train_gold[f'{tag}/global_variables']).split('\n')
  mine = ck.global_variables(p)
  assert sorted(n for n, _ in mine) == sorted(ref_globals) and len(set(ref_globals)) == len(ref_globals)
  ckpt_vars = []
  for line in str(train_gold[f'{tag}/checkpoint_variables']).split('\n'):
    name, *dims = line.split()
    ckpt_vars.append((name, tuple(int(d) for d in dims)))
  ref_map = dict(l.split() for l in str(train_gold[f'{tag}/init_from_checkpoint']).split('\n'))
  assert dict(ck.match_init_checkpoint(ckpt_vars, mine, psp_module=psp)) == ref_map
  assert len(ref_map) == 53 * 5
  # with an init checkpoint the train saver keeps every global variable (exclude = [])
  assert sorted(str(train_gold[f'{tag}/train_saver']).split('\n')) == sorted(ref_globals)


# ------------------------------------------------------------------------------------------------ predict.py
def test_predict_driver_and_exports_equal_the_reference_run(tmp_path):
  """code/predict.py::main executed by the reference (tests/golden/make_reference_predict_fixtures.py): parsed flags,
  `_add_extra_args`, SemanticSegmentation's derived settings, what the estimator is asked for, and the export loop
  (:137-164) on two fixed examples - against wlseg.settings + wlseg.system_factory + wlseg.cli.export_outputs: the same
  settings, the same file names, the same PNG pixels (label ids through cids2lids, colours through cids2colors, the
  50:50 overlay truncated to uint8)."""
  import json
  from PIL import Image
  from wlseg import cli, problem_defs, settings as wsettings
  from wlseg import system_factory as sf
  gold = np.load(os.path.join(os.path.dirname(TRAIN_GOLD), 'reference_predict_run.npz'))
  argv = json.loads(str(gold['argv']))
  problem_defs.write_all()
  argv[0] = problem_defs.default_path('cityscapes')       # the same problem definition, at the product's location
  results = tmp_path / 'results'
  results.mkdir()
  st = wsettings.build_parser(wsettings.PREDICT).parse_args([str(tmp_path / 'log')] + argv + ['--results_dir', str(results)])
  st = wsettings.predict_extra_args(st)
  st.rank, st.world_size = 0, 1
  system = sf.SemanticSegmentation({'predict': lambda config, params: iter(())}, None, st)
  mine = vars(system.settings)
  ref = json.loads(str(gold['settings']))
  for k, v in ref.items():
    if k in _PATH_KEYS:
      continue
    assert k in mine, f'the reference sets settings.{k}, the product does not'
    assert mine[k] == v, f'settings.{k}: {mine[k]!r} != {v!r} (reference)'
  (_, asked), = json.loads(str(gold['calls']))
  assert asked['predict_keys'] == system.settings.predict_keys and asked['checkpoint_path'] == system.settings.ckpt_path
  s = system.settings
  idspalette = np.array(s.inference_problem_def['cids2lids'], dtype=np.uint8)
  colorpalette = np.array(s.inference_problem_def['cids2colors'], dtype=np.uint8)
  for i in range(2):
    cli.export_outputs({'decisions': gold[f'example{i}/decisions'], 'rawimages': gold[f'example{i}/rawimages'],
                        'rawimagespaths': str(gold[f'example{i}/rawimagespaths']).encode()}, s, idspalette, colorpalette)
  names = sorted(os.listdir(str(results)))
  assert names == str(gold['files']).split('\n')
  for n in names:
    got = np.asarray(Image.open(str(results / n)))
    want = gold[f'png/{n}']
    assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), n


def test_evaluate_script_outputs_equal_the_reference_run(tmp_path, monkeypatch):
  """code/evaluate.py::main executed by the reference over a stub estimator (fixed 20 x 20 confusion matrix, step 1234):
  the files it leaves in eval_00/ - `all_metrics.txt` (step + print_metrics_from_confusion_matrix through `printfile`,
  utils/utils.py:385-446) character by character, and the pickled metrics list - against wlseg.cli.evaluate_main on the
  same argv with the same fake estimator."""
  import json
  import pickle
  from wlseg import cli, problem_defs
  from wlseg import system_factory as sf
  gold = np.load(os.path.join(os.path.dirname(TRAIN_GOLD), 'reference_predict_run.npz'))
  argv = json.loads(str(gold['evaluate/argv']))
  problem_defs.write_all()
  argv[1] = problem_defs.default_path('cityscapes')
  log_dir = tmp_path / 'log'
  log_dir.mkdir()
  calls = []

  def fake_create(self, *a, **k):
    self._estimator = _FakeEstimator(self, calls)
  monkeypatch.setattr(sf.SemanticSegmentation, '_create_estimator', fake_create)
  import contextlib
  with contextlib.redirect_stdout(io.StringIO()):
    cli.evaluate_main([str(log_dir)] + argv + ['--synthetic'])
  res = log_dir / 'eval_00'
  assert sorted(os.listdir(str(res))) == str(gold['evaluate/files']).split('\n')
  assert (res / 'all_metrics.txt').read_text() == str(gold['evaluate/all_metrics_txt'])
  with open(str(res / 'all_metrics.p'), 'rb') as fp:
    pickled = pickle.load(fp)
  assert len(pickled) == int(gold['evaluate/pickle_len'])
  assert sorted(pickled[0].keys()) == str(gold['evaluate/pickle_keys']).split('\n')
  assert pickled[0]['global_step'] == int(gold['evaluate/pickle_global_step'])
  assert pickled[0]['confusion_matrix'].dtype == np.int32 and np.array_equal(pickled[0]['confusion_matrix'], gold['evaluate/pickle_cm'])


def test_predict_input_equals_the_reference_pipeline_run(tmp_path):
  """dataset_agnostic_predict_input.py: `_predict_image_generator` (:88-107) + `_predict_preprocess` (:109-119) run by the
  reference on a small directory tree (RGB / grey / palette / RGBA, nested folders, upper-case extension, PPM, a
  non-image file) vs wlseg.image_input.predict_input_fn on the same tree: the same set of files, raw RGB arrays bit-exact,
  processed images (uint8 -> [0, 1] by the float32 constant 1 / 255 -> legacy bilinear resize -> [-1, 1)) bit-exact."""
  import argparse
  import contextlib
  from wlseg import image_input
  gen = importlib.import_module('tests.golden.make_reference_predict_input_fixtures')
  gold = np.load(gen.OUT)
  root = str(tmp_path / 'images')
  os.makedirs(root)
  gen.write_images(root)
  hf, wf = (int(v) for v in gold['size'])
  params = argparse.Namespace(predict_dir=root, height_feature_extractor=hf, width_feature_extractor=wf, Nb=1,
                              preserve_aspect_ratio=False)
  with contextlib.redirect_stdout(io.StringIO()):
    batches = list(image_input.predict_input_fn(None, params))
  seen = []
  worst = 0.0
  for features, labels in batches:
    assert labels is None and sorted(features) == ['proimages', 'rawimages', 'rawimagespaths']
    (path,) = features['rawimagespaths']
    key = os.path.relpath(path.decode('utf-8'), root)
    seen.append(key)
    raw, pro = features['rawimages'], features['proimages']
    assert raw.dtype == torch.uint8 and tuple(raw.shape[:1]) == (1,) and np.array_equal(raw[0].numpy(), gold[f'raw/{key}'])
    assert pro.dtype == torch.float32 and tuple(pro.shape) == (1, hf, wf, 3)
    worst = max(worst, float(np.abs(pro[0].numpy() - gold[f'pro/{key}']).max()))
  assert sorted(seen) == str(gold['paths']).split('\n')
  assert worst == 0.0, worst


@pytest.mark.parametrize('dataset', ['cityscapes', 'vistas'])
def test_regenerated_problem_definitions_equal_the_reference_files(dataset):
  """wlseg/problem_defs.py regenerates code/problem_definitions/<dataset>/problem01.json from the public label tables;
  every field except the free-text `comments` has the SHA-256 of the reference's value
  (tests/golden/make_problem_def_digests.py)."""
  import json
  from wlseg import problem_defs
  gen = importlib.import_module('tests.golden.make_problem_def_digests')
  with open(gen.OUT) as fp:
    want = json.load(fp)[dataset]
  mine = problem_defs.GENERATORS[dataset]()
  assert sorted(k for k in mine if k != 'comments') == sorted(want)
  for k, d in want.items():
    assert gen.digest(mine[k]) == d, f'{dataset}: field {k} differs from the reference file'
