"""Oracle for the hierarchical strong + weak masked cross-entropy losses.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned against the reference-run vectors
(define_losses / _segment_sum executed over tests/golden/tf_shim; tests/test_reference_fixtures.py).

Follows code/estimator/define_losses_hierarchical.py:14-224 step by step;
gradients come from torch autograd on the same expression, which restates the
TF gradient of each op.
"""

import torch

from oracle import tfops
from oracle.tables import TABLES


def l2_targets(per_pixel_labels, per_bbox_labels, per_image_labels, pp2x, bb2x):
  """define_losses_hierarchical.py:110-126: one-hot of the mapped strong label,
  `_segment_sum` of the weak 15-way multinomials, concatenated (strong, bbox, image)."""
  width = max(pp2x) + 1
  assert width == max(bb2x) + 1
  pp = torch.tensor(pp2x, dtype=torch.long)[per_pixel_labels.long()]
  parts = [torch.nn.functional.one_hot(pp, width).to(torch.float32)]
  if per_bbox_labels is not None and per_bbox_labels.shape[0] > 0:
    parts.append(tfops.unsorted_segment_sum_last(per_bbox_labels, bb2x, width))
  if per_image_labels is not None and per_image_labels.shape[0] > 0:
    parts.append(tfops.unsorted_segment_sum_last(per_image_labels, bb2x, width))
  return torch.cat(parts, 0)


def l2_weights(targets, l1_decisions, n_strong, cid_l1):
  """define_losses_hierarchical.py:154-165 (vehicle) / :175-185 (human)."""
  strong_w = 1.0 - targets[:n_strong, ..., -1]
  weak_t = targets[n_strong:]
  not_void = (1.0 - weak_t[..., -1]) > 0.01
  l1_correct = (l1_decisions[n_strong:] == cid_l1) & (weak_t[..., :-1].max(dim=-1).values >= 0.01)
  weak_w = (not_void & l1_correct).to(torch.float32)
  return torch.cat([strong_w, weak_w], 0)


def define_losses(predictions, labels, dataset, conv_weights=(), regularization_weight=0.0):
  """TRAIN branch of define_losses.  `predictions` holds full-resolution
  l1/l2_vehicle/l2_human logits and 'l1_decisions'; `labels` the three
  prolabels_* tensors (missing / empty weak parts allowed).
  Returns the reference's loss dict plus the three nonzero-weight counts."""
  t = TABLES[dataset]
  pp = labels['prolabels_per_pixel']
  pb = labels.get('prolabels_per_bbox')
  pi = labels.get('prolabels_per_image')
  n_strong = pp.shape[0]

  # L1: sparse CE on the strong part only (:131-135, :191-192)
  pp2l1 = torch.tensor(t['per_pixel_cids2l1_cids'], dtype=torch.long)
  y1 = pp2l1[pp.long()]
  ce1 = tfops.sparse_softmax_cross_entropy(predictions['l1_logits'][:n_strong], y1)
  w1 = (y1 <= max(t['per_pixel_cids2l1_cids']) - 1).to(torch.float32)
  l1_loss, n1 = tfops.compute_weighted_loss(ce1, w1)

  l1_decs = predictions['l1_decisions']
  out = {}
  counts = {'l1': n1}
  for head, pp2x, bb2x, cid in (
      ('l2_vehicle', t['per_pixel_cids2vehicle_cids'], t['per_bbox_cids2vehicle_cids'], t['cid_l1_vehicle']),
      ('l2_human', t['per_pixel_cids2human_cids'], t['per_bbox_cids2human_cids'], t['cid_l1_human'])):
    tg = l2_targets(pp, pb, pi, pp2x, bb2x)
    ce = tfops.softmax_cross_entropy(predictions[f'{head}_logits'], tg)
    w = l2_weights(tg, l1_decs, n_strong, cid)
    out[head], counts[head] = tfops.compute_weighted_loss(ce, w)

  seg = l1_loss + 0.1 * (out['l2_vehicle'] + out['l2_human'])  # :200-204
  reg = torch.zeros(())
  for w in conv_weights:  # slim.l2_regularizer: s * sum(w^2) / 2
    reg = reg + regularization_weight * 0.5 * (w ** 2).sum()
  return {'total': seg + reg,
          'segmentation': seg,
          'l1_segmentation': l1_loss,
          'l1_segmentation_hot': torch.zeros(()),
          'l2_vehicle_segmentation': out['l2_vehicle'],
          'l2_human_segmentation': out['l2_human'],
          'regularization': reg,
          'counts': counts}
