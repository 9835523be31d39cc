"""Oracle for the integer evaluation path: cid remap, confusion matrix, metrics.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Integer work is numpy;
`confmat_oracle.c` is the same histogram in plain C (built by oracle/Makefile).
"""

import ctypes
import os

import numpy as np


def replacevoids(mappings):
  """code/utils/utils.py:286-289: -1 -> max+1."""
  max_m = max(mappings)
  return [m if m != -1 else max_m + 1 for m in mappings]


def map_decisions_to_new_cids(decisions, old_cids2new_cids):
  """code/estimator/define_estimator_hierarchical.py:511-514 (decisions part)."""
  lut = np.asarray(replacevoids(list(old_cids2new_cids)), dtype=np.int32)
  return lut[np.asarray(decisions)]


def map_probabilities_to_new_cids(probs, old_cids2new_cids):
  """:515-519: probabilities of merged classes are summed (worked example :494-496)."""
  lut = np.asarray(replacevoids(list(old_cids2new_cids)), dtype=np.int64)
  out = np.zeros(probs.shape[:-1] + (int(lut.max()) + 1,), dtype=probs.dtype)
  for old, new in enumerate(lut):
    out[..., new] += probs[..., old]
  return out


def replace_voids_hierarchical(l1_probs, l2v_probs, l2h_probs, decisions, tables):
  """`_replace_voids` (code/estimator/define_estimator_hierarchical.py:573-630) carried over to the
  hierarchical classifier.  The reference takes tf.nn.top_k(probs, 2) of a flat classifier whose LAST
  channel is void and, where the decision is void, the runner-up `indices[..., 1]` (:611-622); on this
  model it stops at its own key-set assert (:589-592: the l2_human_* keys are not in `supported_keys`).
  Per head the same rule reads: a head that chose its void channel (its last one) takes its best non-void
  class instead; the pixel's decision is then composed again
  (code/models/resnet50_extended_model_hierarchical.py:95-117).  Only void decisions change.
  `tables` is oracle.tables.TABLES[dataset]."""
  p1, pv, ph = (np.asarray(a, dtype=np.float32) for a in (l1_probs, l2v_probs, l2h_probs))
  decs = np.asarray(decisions).astype(np.int32).copy()
  l1_to_common = np.asarray(tables['l1_cids2common_cids'], dtype=np.int32)
  veh_to_common = np.asarray(tables['l2_vehicle_cids2common_cids'], dtype=np.int32)
  hum_to_common = np.asarray(tables['l2_human_cids2common_cids'], dtype=np.int32)
  void_cid = int(l1_to_common[-1])
  d1 = np.argmax(p1[..., :-1], -1)   # np.argmax: first maximum, as tf.argmax / top_k
  dv = np.argmax(pv[..., :-1], -1)
  dh = np.argmax(ph[..., :-1], -1)
  new = np.where(d1 == tables['cid_l1_vehicle'], veh_to_common[dv],
                 np.where(d1 == tables['cid_l1_human'], hum_to_common[dh], l1_to_common[d1]))
  return np.where(decs == void_cid, new, decs).astype(np.int32)


def confusion_matrix(labels, decisions, num_classes):
  """[TF-1.12] metrics_impl._streaming_confusion_matrix update for one batch:
  cm[label, prediction] += 1 over all pixels
  (code/estimator/define_estimator_hierarchical.py:185-194).  int64 result."""
  lab = np.asarray(labels).reshape(-1).astype(np.int64)
  dec = np.asarray(decisions).reshape(-1).astype(np.int64)
  cm = np.zeros((num_classes, num_classes), dtype=np.int64)
  np.add.at(cm, (lab, dec), 1)
  return cm


_LIB = None


def _clib():
  global _LIB
  if _LIB is None:
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'libconfmat_oracle.so')
    _LIB = ctypes.CDLL(path)
    _LIB.oracle_confusion_matrix.restype = ctypes.c_int
    _LIB.oracle_confusion_matrix.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
    _LIB.oracle_compose_decisions.restype = None
  return _LIB


def confusion_matrix_c(labels, decisions, num_classes):
  """Same histogram through the plain-C restatement (oracle/confmat_oracle.c)."""
  lab = np.ascontiguousarray(np.asarray(labels).reshape(-1), dtype=np.int32)
  dec = np.ascontiguousarray(np.asarray(decisions).reshape(-1), dtype=np.int32)
  cm = np.zeros((num_classes, num_classes), dtype=np.int64)
  bad = _clib().oracle_confusion_matrix(
      lab.ctypes.data, dec.ctypes.data, lab.size, num_classes, cm.ctypes.data)
  assert bad == 0, f'{bad} out-of-range (label, decision) pairs'
  return cm


def metrics_from_confusion_matrix(cm):
  """code/utils/utils.py:414-423.  Returns dict with percentages as the reference prints."""
  cm = np.asarray(cm)
  with np.errstate(divide='ignore', invalid='ignore'):
    global_accuracy = np.trace(cm) / np.sum(cm) * 100
    accuracies = np.diagonal(cm) / np.sum(cm, 1) * 100
    inter = np.diagonal(cm)
    union = np.sum(cm, 0) + np.sum(cm, 1) - np.diagonal(cm)
    ious = inter / np.where(union > 0, union, np.ones_like(union)) * 100
  notnan = np.logical_not(np.isnan(accuracies))
  return {'global_accuracy': global_accuracy,
          'accuracies': accuracies,
          'ious': ious,
          'notnan_mask': notnan,
          'mean_accuracy': np.mean(accuracies[notnan]),
          'mean_iou': np.mean(ious[notnan])}


def batch_mean_iou(labels, decisions, num_classes):
  """code/estimator/define_metrics.py:5-20 (training summary): int32 cm ->
  mean over ALL classes of inter / (union + 1e-9), fp32."""
  cm = confusion_matrix(labels, decisions, num_classes).astype(np.int32)
  inter = np.diagonal(cm).astype(np.float32)
  union = (cm.sum(0) + cm.sum(1) - np.diagonal(cm)).astype(np.float32) + np.float32(1e-9)
  return np.mean(inter / union, dtype=np.float32)
