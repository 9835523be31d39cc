"""Pins the oracle against every worked example / known answer the reference carries for this
path.  The reference has no executable tests for the model, loss or
metrics, so these (plus published TF-1.12 op semantics) are all there is: PARITY UNPINNED."""

import json
import os

import numpy as np
import torch

from oracle import losses as olosses
from oracle import metrics as ometrics
from oracle import network as onet
from oracle import optimizer as oopt
from oracle import tfops
from oracle import weak_labels as oweak
from oracle.tables import TABLES

HERE = os.path.dirname(os.path.abspath(__file__))


def test_cid_remap_worked_example():
  # code/estimator/define_estimator_hierarchical.py:494-498
  assert ometrics.replacevoids([-1, 1, 1, 0, -1]) == [2, 1, 1, 0, 2]
  decs = np.array([[0, 1, 2, 3, 4]])
  assert ometrics.map_decisions_to_new_cids(decs, [-1, 1, 1, 0, -1]).tolist() == [[2, 1, 1, 0, 2]]
  probs = np.array([[0.1, 0.2, 0.3, 0.15, 0.25]], dtype=np.float32)
  out = ometrics.map_probabilities_to_new_cids(probs, [-1, 1, 1, 0, -1])
  assert out.shape == (1, 3)
  assert np.allclose(out, [[0.15, 0.5, 0.35]])


def test_bbox_label_normalisation_examples():
  # code/input_pipelines/open_images/input_subset_bboxes_v2.py:87-95: [car, bus, person] counts
  full = (0.0, 0.999, 0.0, 0.999)
  car, bus, human = 2, 1, 6
  lab = oweak.bbox_labels([(car, *full)], 4, 4)
  assert lab[0, 0, car] == 1.0
  lab = oweak.bbox_labels([(car, *full), (car, *full)], 4, 4)
  assert lab[0, 0, car] == 1.0
  lab = oweak.bbox_labels([(car, *full), (bus, *full)], 4, 4)
  assert np.allclose(lab[0, 0, [car, bus]], [0.5, 0.5])
  lab = oweak.bbox_labels([(car, *full), (car, *full), (bus, *full)], 4, 4)
  assert np.allclose(lab[0, 0, [car, bus]], [2 / 3, 1 / 3])
  lab = oweak.bbox_labels([(human, 0.0, 0.4, 0.0, 0.4)], 8, 8)
  assert lab[7, 7, 14] == 1.0 and lab[7, 7, :14].sum() == 0  # no box -> void
  # the only numeric assertion in the reference: every pixel sums to 1 (+-1e-3 / 1e-2),
  # input_subset_bboxes_v2_test.py:40-43, input_subset_image_labels_test.py:41-44
  rng = np.random.default_rng(0)
  boxes = [(int(rng.integers(0, 14)), *sorted(rng.random(2) * 0.99), *sorted(rng.random(2) * 0.99)) for _ in range(12)]
  assert np.all(np.abs(oweak.bbox_labels(boxes, 33, 47).sum(-1) - 1) < 1e-3)
  assert np.all(np.abs(oweak.image_labels([0, 5, 9], 5, 6).sum(-1) - 1) < 1e-2)
  assert oweak.image_labels([], 2, 2)[0, 0, 14] == 1.0


def test_segment_sum_example():
  # define_losses_hierarchical.py:112-113: human box + vehicle box -> vehicle head sees 1/2 + 1/2 void
  t = TABLES['cityscapes']
  weak = torch.zeros(1, 1, 1, 15)
  weak[..., 2] = 0.5  # car
  weak[..., 6] = 0.5  # human
  veh = tfops.unsorted_segment_sum_last(weak, t['per_bbox_cids2vehicle_cids'], 7)
  assert veh[0, 0, 0].tolist() == [0.5, 0, 0, 0, 0, 0, 0.5]
  hum = tfops.unsorted_segment_sum_last(weak, t['per_bbox_cids2human_cids'], 3)
  assert hum[0, 0, 0].tolist() == [0.5, 0, 0.5]


def test_output_nclasses_and_cm_trim():
  # system_factory.py:124-130,400-405 with the cityscapes problem definition
  lids2cids = [-1, -1, -1, -1, -1, -1, -1, 0, 1, -1, -1, 2, 3, 4, -1, -1, -1, 5, -1, 6, 7, 8, 9, 10, 11, 12, 13, 14,
               15, -1, -1, 16, 17, 18]
  output_nclasses = max(lids2cids) + 1 + (-1 in lids2cids)
  assert output_nclasses == 20
  tcids2ecids = list(range(output_nclasses))
  tcids2ecids[-1] = -1
  assert max(ometrics.replacevoids(tcids2ecids)) + 1 == 20
  cm = np.arange(400).reshape(20, 20)
  assert cm[:-1, :-1].shape == (19, 19)


def test_hierarchy_round_trip_and_head_widths():
  for ds, t in TABLES.items():
    n = t['num_classes']
    assert len(t['per_pixel_cids2l1_cids']) == n
    assert t['head_widths'] == (max(t['per_pixel_cids2l1_cids']) + 1, max(t['per_pixel_cids2vehicle_cids']) + 1,
                                max(t['per_pixel_cids2human_cids']) + 1)
    for cid in range(n):
      l1 = t['per_pixel_cids2l1_cids'][cid]
      if l1 == t['cid_l1_vehicle']:
        back = t['l2_vehicle_cids2common_cids'][t['per_pixel_cids2vehicle_cids'][cid]]
      elif l1 == t['cid_l1_human']:
        back = t['l2_human_cids2common_cids'][t['per_pixel_cids2human_cids'][cid]]
      else:
        back = t['l1_cids2common_cids'][l1]
      assert back == cid, (ds, cid)


def test_lr_schedule_defaults():
  # system_factory.py:207-233 with train.py defaults (Cityscapes)
  s = oopt.train_schedule()
  assert s['num_batches_per_epoch'] == 743 and s['num_training_steps'] == 17 * 743
  assert s['boundaries'] == [8 * 743, 15 * 743]
  assert s['values'] == [0.01, 0.005, 0.0025]
  assert oopt.piecewise_constant(0, s['boundaries'], s['values']) == 0.01
  assert oopt.piecewise_constant(8 * 743, s['boundaries'], s['values']) == 0.01
  assert oopt.piecewise_constant(8 * 743 + 1, s['boundaries'], s['values']) == 0.005
  assert oopt.piecewise_constant(10 ** 6, s['boundaries'], s['values']) == 0.0025
  assert abs(oopt.polynomial_decay(0.01, 0, 100, 0.5, 0.9) - 0.01) < 1e-12
  assert abs(oopt.polynomial_decay(0.01, 100, 100, 0.5, 0.9) - 0.5) < 1e-12


def test_metrics_from_confusion_matrix():
  # code/utils/utils.py:414-423: union 0 -> IoU 0 but excluded (empty row); means over non-NaN rows
  cm = np.array([[3, 1, 0], [2, 4, 0], [0, 0, 0]], dtype=np.int32)
  m = ometrics.metrics_from_confusion_matrix(cm)
  assert np.isclose(m['global_accuracy'], 70.0)
  assert np.allclose(m['accuracies'][:2], [75.0, 400 / 6])
  assert np.isnan(m['accuracies'][2]) and m['ious'][2] == 0.0
  assert np.allclose(m['ious'][:2], [50.0, 400 / 7])
  assert np.isclose(m['mean_accuracy'], np.mean([75.0, 400 / 6]))
  assert np.isclose(m['mean_iou'], np.mean([50.0, 400 / 7]))
  # define_metrics.py: mean over ALL classes of inter / (union + 1e-9)
  lab = np.array([0, 0, 0, 0, 1, 1, 1, 1, 1, 1])
  dec = np.array([0, 0, 0, 1, 0, 0, 1, 1, 1, 1])
  assert np.isclose(ometrics.batch_mean_iou(lab, dec, 3), np.mean([3 / 6, 4 / 7, 0.0]), atol=1e-6)


def test_tf_same_padding_and_resize_semantics():
  # pool1 3x3/2 on even input pads (0,1); on odd input (1,1)
  assert tfops.same_pad(8, 3, 2) == (0, 1, 4)
  assert tfops.same_pad(9, 3, 2) == (1, 1, 5)
  assert tfops.same_pad(10, 3, 1, rate=4) == (4, 4, 10)
  x = torch.arange(16, dtype=torch.float32).view(1, 4, 4, 1)
  y = tfops.max_pool_same(x, 3, 2)
  assert y.view(-1).tolist() == [10.0, 11.0, 14.0, 15.0]
  # align_corners bilinear: corners preserved, midpoints interpolated
  z = tfops.resize_bilinear(torch.tensor([0.0, 1.0, 2.0]).view(1, 1, 3, 1), 1, 5)
  assert torch.allclose(z.view(-1), torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0]))
  n = tfops.resize_nearest(torch.tensor([0, 1, 2]).view(1, 1, 3), 1, 5)
  assert n.view(-1).tolist() == [0, 1, 1, 2, 2]  # roundf(0.5) = 1, roundf(1.5) = 2
  assert tfops.argmax_first(torch.tensor([[1.0, 3.0, 3.0, 2.0]])).tolist() == [1]


def test_network_shapes_and_parameter_count():
  specs = onet.conv_specs('cityscapes')
  assert len(specs) == 66
  assert sum(a * b * c * d for a, b, c, d in specs.values()) == 26148032  # SURVEY: 26.15 M conv weights
  p = onet.init_params('cityscapes', seed=0)
  out = onet.Net(p, 'cityscapes').forward(torch.zeros(1, 32, 64, 3))
  assert [tuple(z.shape) for z in out['lowres_logits']] == [(1, 4, 8, 14), (1, 4, 8, 7), (1, 4, 8, 3)]
  assert out['decisions'].shape == (1, 32, 64) and out['decisions'].dtype == torch.int32
  assert set(out) >= {'l1_logits', 'l1_probabilities', 'l1_decisions', 'l2_vehicle_logits',
                      'l2_vehicle_probabilities', 'l2_vehicle_decisions', 'l2_human_logits',
                      'l2_human_probabilities', 'l2_human_decisions', 'decisions'}
  v = onet.conv_specs('vistas')
  assert [v[f'softmax_classifier/{k}'][3] for k in ('l1_logits', 'l2_vehicle_logits', 'l2_human_logits')] == [53, 12, 5]


def test_loss_weights_and_reduction_semantics():
  """SUM_BY_NONZERO_WEIGHTS with safe-div; L1 ignores void; weak pixels count only where the L1
  argmax is the right super-class (define_losses_hierarchical.py:142-187)."""
  t = TABLES['cityscapes']
  H = W = 2
  l1 = torch.zeros(2, H, W, 14)
  l1[1, ..., t['cid_l1_vehicle']] = 4.0          # weak image: L1 says vehicle everywhere
  l1[1, 0, 0, t['cid_l1_human']] = 9.0           # ... except one pixel
  pred = {'l1_logits': l1, 'l2_vehicle_logits': torch.zeros(2, H, W, 7), 'l2_human_logits': torch.zeros(2, H, W, 3),
          'l1_decisions': tfops.argmax_first(l1)}
  strong = torch.tensor([[[0, 19], [13, 11]]], dtype=torch.int32)  # road, void, car, person
  bbox = torch.from_numpy(oweak.bbox_labels([(2, 0.0, 0.999, 0.0, 0.999)], H, W))[None]
  out = olosses.define_losses(pred, {'prolabels_per_pixel': strong, 'prolabels_per_bbox': bbox}, 'cityscapes')
  assert out['counts']['l1'] == 3                 # void dropped
  assert out['counts']['l2_vehicle'] == 1 + 3     # strong car + 3 weak pixels with L1 == vehicle
  assert out['counts']['l2_human'] == 1           # strong person only (weak target has no human mass)
  assert np.isclose(float(out['l1_segmentation']), np.log(14.0), atol=1e-6)
  assert np.isclose(float(out['l2_vehicle_segmentation']), np.log(7.0), atol=1e-6)
  assert np.isclose(float(out['segmentation']), np.log(14) + 0.1 * (np.log(7) + np.log(3)), atol=1e-6)
  none = olosses.define_losses(pred, {'prolabels_per_pixel': torch.full((1, H, W), 19, dtype=torch.int32)}, 'cityscapes')
  assert float(none['l1_segmentation']) == 0.0   # safe-div: no nonzero weight -> 0


def test_golden_fixture_matches_oracle():
  """tests/golden/oracle_small.json was written by tests/golden/make_golden.py from this oracle;
  it freezes the oracle's numbers so that later edits to the oracle cannot drift silently."""
  path = os.path.join(HERE, 'golden', 'oracle_small.json')
  with open(path) as fp:
    gold = json.load(fp)
  from tests.golden import make_golden
  now = make_golden.compute()
  for key, val in gold.items():
    a, b = np.asarray(val, dtype=np.float64), np.asarray(now[key], dtype=np.float64)
    assert a.shape == b.shape, key
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6), key


def test_cross_replica_batch_norm_restatement():
  """--cross_replica_norm (utils/cross_replica_batch_normalization.py:398-459): with equal per-replica
  batch sizes the global moments are those of the concatenated batch, so the outputs equal ordinary
  batch norm over the concatenation; the moving variance takes the BIASED global variance times
  (n_local - 1) / n_local - not Bessel's correction over the global sample."""
  import torch
  from oracle import tfops
  g = torch.Generator().manual_seed(2)
  xs = [torch.randn(2, 3, 5, 4, generator=g) * (r + 1) + r for r in range(2)]
  gamma, beta = torch.rand(4, generator=g) + 0.5, torch.randn(4, generator=g)
  mm, mv = torch.randn(4, generator=g), torch.rand(4, generator=g) + 0.5
  ys, new_mm, new_mv, mean, var = tfops.cross_replica_batch_norm(xs, gamma, beta, mm, mv, decay=0.9, eps=1e-5)
  cat = torch.cat(xs, 0)
  y_ref, ref_mm, ref_mv, ref_mean, ref_var = tfops.batch_norm(cat, gamma, beta, mm, mv, True, decay=0.9, eps=1e-5)
  assert torch.allclose(torch.cat(ys, 0), y_ref, rtol=1e-5, atol=1e-5)
  assert torch.allclose(mean, ref_mean, atol=1e-6) and torch.allclose(var, ref_var, rtol=1e-5, atol=1e-6)
  assert torch.allclose(new_mm, ref_mm, atol=1e-6)
  n_local = 2 * 3 * 5
  assert torch.allclose(new_mv, mv - 0.1 * (mv - ref_var * (n_local - 1) / n_local), rtol=1e-5, atol=1e-6)
  assert not torch.allclose(new_mv, ref_mv, rtol=1e-4)  # it is NOT the single-process update


def test_replace_voids_hierarchical_known_answers():
  """`_replace_voids` (define_estimator_hierarchical.py:573-630) per head on hand-made probabilities:
  only void decisions change; an L1 void falls to the L1 runner-up (composed through L2 when that is a
  super-class); an L2 void under a vehicle / human decision falls to that head's runner-up."""
  import numpy as np
  from oracle import metrics as ometrics
  from oracle.tables import TABLES
  t = TABLES['cityscapes']
  C1, Cv, Ch = t['head_widths']

  def onehotish(c, idx, second=None):
    p = np.full(c, 0.01, np.float32)
    p[idx] = 0.6
    if second is not None:
      p[second] = 0.3
    return p / p.sum()
  # pixel 0: road (0), not void -> unchanged.  pixel 1: L1 void, runner-up sky-ish class 10 -> 10.
  # pixel 2: L1 void, runner-up vehicle (12), vehicle head best non-void = 2 -> common 15.
  # pixel 3: L1 vehicle, vehicle head void (6), runner-up 4 -> common 17.
  # pixel 4: L1 human (11), human head void (2), runner-up 1 -> common 12 (rider).
  p1 = np.stack([onehotish(C1, 0), onehotish(C1, 13, 10), onehotish(C1, 13, 12), onehotish(C1, 12), onehotish(C1, 11)])
  pv = np.stack([onehotish(Cv, 6), onehotish(Cv, 6), onehotish(Cv, 6, 2), onehotish(Cv, 6, 4), onehotish(Cv, 0)])
  ph = np.stack([onehotish(Ch, 2), onehotish(Ch, 2), onehotish(Ch, 2), onehotish(Ch, 0), onehotish(Ch, 2, 1)])
  decs = np.array([0, 19, 19, 19, 19], np.int32)
  got = ometrics.replace_voids_hierarchical(p1[None, None], pv[None, None], ph[None, None], decs[None, None], t)
  assert got.reshape(-1).tolist() == [0, 10, 15, 17, 12]


def test_resize_nearest_is_roundf():
  """[TF-1.12] ResizeNearestNeighbor, align_corners: src = min(roundf(dst * scale), in - 1); a 4 -> 7 resize
  has scale 0.5: dst 1 -> roundf(0.5) = 1 (half away from zero), dst 3 -> roundf(1.5) = 2."""
  import torch
  from oracle import tfops
  x = torch.arange(4, dtype=torch.int32).view(1, 4, 1)
  got = tfops.resize_nearest(x, 7, 1, align_corners=True).reshape(-1).tolist()
  assert got == [0, 1, 1, 2, 2, 3, 3]


def test_new_oracle_ops_against_independent_formulations():
  """The oracle functions added for the optional model parts, each against a second, independent statement of the
  same [TF-1.12] op: group_norm vs torch's own group norm on the channel-permuted tensor; conv2d_transpose (stride
  1, SAME) vs the explicit double loop over taps; VALID average pooling vs window means; align_corners bilinear
  resize at a hand-computable size."""
  import torch
  import torch.nn.functional as F
  from oracle import tfops
  g = torch.Generator().manual_seed(0)
  # group norm: [N, H, W, C] with contiguous channel groups == F.group_norm on NCHW
  x = torch.randn(2, 5, 4, 64, generator=g)
  gamma, beta = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
  want = F.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5).permute(0, 2, 3, 1)
  assert torch.allclose(tfops.group_norm(x, gamma, beta, 32, 1e-5), want, atol=1e-5)
  want1 = F.group_norm(x.permute(0, 3, 1, 2), 1, gamma, beta, eps=1e-5).permute(0, 2, 3, 1)   # logits layers: 1 group
  assert torch.allclose(tfops.group_norm(x, gamma, beta, 1, 1e-5), want1, atol=1e-5)
  # conv2d_transpose, filter [kh, kw, out, in]: out[y, x, o] = sum_{r, s, i} in[y + 1 - r, x + 1 - s, i] * f[r, s, o, i]
  xin = torch.randn(1, 4, 5, 3, generator=g)
  f = torch.randn(3, 3, 2, 3, generator=g)
  b = torch.randn(2, generator=g)
  want = torch.zeros(1, 4, 5, 2)
  for y in range(4):
    for xx in range(5):
      for r in range(3):
        for s in range(3):
          yy, xs = y + 1 - r, xx + 1 - s
          if 0 <= yy < 4 and 0 <= xs < 5:
            want[0, y, xx] += f[r, s] @ xin[0, yy, xs]
  assert torch.allclose(tfops.conv2d_transpose_same(xin, f, b), want + b, atol=1e-5)
  # VALID average pooling: windows that do not fit are dropped
  xp = torch.arange(2 * 5 * 7 * 1, dtype=torch.float32).view(2, 5, 7, 1)
  got = tfops.avg_pool_valid(xp, (2, 3), (2, 3))
  assert tuple(got.shape) == (2, 2, 2, 1)
  assert float(got[1, 1, 1, 0]) == float(xp[1, 2:4, 3:6, 0].mean())
  # align_corners bilinear 2 -> 3: the middle sample is the exact midpoint, corners are kept
  xr = torch.tensor([[[[0.0], [4.0]], [[8.0], [12.0]]]])
  got = tfops.resize_bilinear(xr, 3, 3, align_corners=True)[0, :, :, 0]
  assert got.tolist() == [[0.0, 2.0, 4.0], [4.0, 6.0, 8.0], [8.0, 10.0, 12.0]]


def test_product_convolution_geometry_equals_the_oracle_padding_rules():
  """wlseg/arch.py::same_pad_before (leading zero padding + output size of every slim convolution: TF 'SAME' for
  stride 1 and for 1x1 kernels, resnet_utils.conv2d_same's explicit padding for strided k > 1) against the oracle's
  tfops.same_pad / conv2d_same arithmetic - which the reference's model() run pins at sizes that are no multiple of 8
  (tests/golden/reference_model_run.npz: vistas_odd_size, cs_odd_size_train_bn) - for every size 1..80."""
  import torch
  from oracle import tfops
  from wlseg import arch
  for size in range(1, 81):
    for k in (1, 3, 7):
      for stride in (1, 2):
        for rate in (1, 2, 4):
          if stride > 1 and rate > 1:
            continue          # never combined: a strided unit past the target stride runs at stride 1 with a rate
          before, out = arch.same_pad_before(k, stride, rate, size)
          if stride == 1 or k == 1:
            want_before, _, want_out = tfops.same_pad(size, k, stride, rate)
          else:
            k_eff = k + (k - 1) * (rate - 1)
            want_before = (k_eff - 1) // 2
            x = torch.zeros(1, size, size, 1)
            want_out = tfops.conv2d_same(x, torch.zeros(k, k, 1, 1), stride, rate).shape[1]
          assert (before, out) == (want_before, want_out), (size, k, stride, rate)
