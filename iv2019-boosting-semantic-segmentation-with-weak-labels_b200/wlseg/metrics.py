"""Metrics from an integer confusion matrix, with the reference's conventions.

Mirrors code/utils/utils.py:385-446 (`print_metrics_from_confusion_matrix`): same input contract
(square np.int32 matrix, asserted), same definitions (global accuracy = trace / sum; class accuracy =
diag / row sum, NaN for empty rows; IoU = diag / (row + col - diag) with a zero union replaced by 1;
means over classes whose accuracy is not NaN) and the same report layout.
"""

import numpy as np


def compute_metrics(cm):
  with np.errstate(divide='ignore', invalid='ignore'):
    diag = np.diagonal(cm)
    rows, cols = np.sum(cm, 1), np.sum(cm, 0)
    global_accuracy = np.trace(cm) / np.sum(cm) * 100
    accuracies = diag / rows * 100
    union = cols + rows - diag
    ious = diag / np.where(union > 0, union, np.ones_like(union)) * 100
  keep = np.logical_not(np.isnan(accuracies))
  return {'global_accuracy': global_accuracy, 'accuracies': accuracies, 'ious': ious, 'notnan_mask': keep,
          'mean_accuracy': np.mean(accuracies[keep]), 'mean_iou': np.mean(ious[keep])}


def print_metrics_from_confusion_matrix(cm, labels=None, printfile=None, printcmd=False, summary=False):
  assert isinstance(cm, np.ndarray), 'Confusion matrix must be numpy array.'
  cms = cm.shape
  assert all([cm.dtype == np.int32, cm.ndim == 2, cms[0] == cms[1], not np.any(np.isnan(cm))]), (
      f"Check print_metrics_from_confusion_matrix input requirements. "
      f"Input has {cm.ndim} dims, is {cm.dtype}, has shape {cms[0]}x{cms[1]} and may contain NaNs.")
  if not labels:
    labels = ['unknown'] * cms[0]
  assert len(labels) == cms[0], (
      f"labels ({len(labels)}) must be enough for indexing confusion matrix ({cms[0]}x{cms[1]}).")
  m = compute_metrics(cm)
  lines = ['', f"Global accuracy: {m['global_accuracy']:5.2f}",
           'Per class accuracies (nans due to 0 #Trues) and ious (nans due to 0 #TPs):']
  # the reference builds a dict keyed by label name first (utils/utils.py:431): rows that share a name - the
  # default 'unknown' labels - collapse into ONE line carrying the last row's values; kept byte for byte
  rows = {name: (acc, iou, ok) for name, acc, iou, ok in zip(labels, m['accuracies'], m['ious'], m['notnan_mask'])}
  for name, (acc, iou, ok) in rows.items():
    lines.append(f"{name:<30s}  {acc:>5.2f}  {iou:>5.2f}  {'' if ok else '(ignored in averages)'}")
  lines.append(f"Mean accuracy (ignoring nans): {m['mean_accuracy']:5.2f}")
  lines.append(f"Mean iou (ignoring accuracies' nans but including ious' 0s): {m['mean_iou']:5.2f}")
  log_string = '\n'.join(lines) + '\n'
  if printcmd:
    print(log_string)
  if printfile:
    if summary:
      printfile.write(log_string)
    else:
      print(f"{m['global_accuracy']:>5.2f}", f"{m['mean_accuracy']:>5.2f}", f"{m['mean_iou']:>5.2f}",
            m['accuracies'], m['ious'], file=printfile)
  return m
