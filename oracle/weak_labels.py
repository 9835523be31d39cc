"""Oracle for the weak-label tensors the (replaced) input side hands to the loss.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  These are numpy
restatements of the two `_generate_rla` functions whose output contract the loss
depends on.
"""

import numpy as np

NUM_WEAK = 15  # 14 Open Images classes + void, input_subset_bboxes_v2.py:38-53


def bbox_labels(boxes, h, w):
  """code/input_pipelines/open_images/input_subset_bboxes_v2.py:74-98.

  boxes: iterable of (cid, xmin, xmax, ymin, ymax) with normalised coordinates.
  Overlapping boxes add counts, every pixel is normalised to a multinomial,
  pixels under no box become void (channel 14) = 1.
  """
  rla = np.zeros((h, w, NUM_WEAK), dtype=np.float32)
  for cid, xmin, xmax, ymin, ymax in boxes:
    x0, x1, y0, y1 = int(xmin * w), int(xmax * w), int(ymin * h), int(ymax * h)
    rla[y0:y1 + 1, x0:x1 + 1, cid] += 1
  s = rla.sum(axis=2, keepdims=True)
  void = np.concatenate([np.zeros(NUM_WEAK - 1, np.float32), np.ones(1, np.float32)])
  with np.errstate(divide='ignore', invalid='ignore'):
    return np.where(s > 0.5, rla / s, void).astype(np.float32)


def image_labels(cids, h, w):
  """code/input_pipelines/open_images/input_subset_image_labels.py:73-107:
  uniform over the present classes (void if none), tiled over the image."""
  v = np.zeros(NUM_WEAK, dtype=np.float32)
  for c in cids:
    v[c] = 1.0
  if not len(cids):
    v[-1] = 1.0
  v /= v.sum()
  return np.tile(v[None, None, :], (h, w, 1))
