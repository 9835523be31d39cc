"""Oracle for the input-side resize + random crop.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Follows code/input_pipelines/utils.py:181-247
(`resize_images_and_labels`) and code/utils/utils.py:540-605 (`resize_images_or_labels`, mode 'max') with the
[TF-1.12] semantics of tf.image.resize_images(align_corners=False): scale = in / out (float32), src = dst * scale,
bilinear lo = floor(src), hi = min(lo + 1, in - 1), x first then y; nearest src = min(floor(dst * scale), in - 1).
Pinned by tests/golden/reference_run.npz (the reference's own function run over the TF shim).
"""

import math

import numpy as np
import torch


def _scale(i, o):
  return np.float32(i) / np.float32(o)


def resize_bilinear_legacy(x, oh, ow):
  ih, iw = x.shape[1], x.shape[2]
  ys = np.arange(oh, dtype=np.float32) * _scale(ih, oh)
  xs = np.arange(ow, dtype=np.float32) * _scale(iw, ow)
  y0 = np.minimum(np.floor(ys).astype(np.int64), ih - 1)
  x0 = np.minimum(np.floor(xs).astype(np.int64), iw - 1)
  y1, x1 = np.minimum(y0 + 1, ih - 1), np.minimum(x0 + 1, iw - 1)
  ly = torch.from_numpy((ys - np.floor(ys)).astype(np.float32)).reshape(1, -1, 1, 1)
  lx = torch.from_numpy((xs - np.floor(xs)).astype(np.float32)).reshape(1, 1, -1, 1)
  x = x.to(torch.float32)
  top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * lx
  bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * lx
  return top + (bot - top) * ly


def resize_nearest_legacy(x, oh, ow):
  ih, iw = x.shape[1], x.shape[2]
  yi = np.minimum(np.floor(np.arange(oh, dtype=np.float32) * _scale(ih, oh)).astype(np.int64), ih - 1)
  xi = np.minimum(np.floor(np.arange(ow, dtype=np.float32) * _scale(iw, ow)).astype(np.int64), iw - 1)
  return x[:, yi][:, :, xi]


def resize_images_and_labels(images, labels, target_size, preserve_aspect_ratio=False, offset=(0, 0)):
  H, W = images.shape[1], images.shape[2]
  th, tw = target_size
  if preserve_aspect_ratio:
    factor = max(th / H, tw / W)                       # float64, the reference's implicit cast (utils.py:578-583)
    RH, RW = int(math.ceil(factor * H)), int(math.ceil(factor * W))
  else:
    RH, RW = th, tw
  pro = resize_bilinear_legacy(images, RH, RW)
  lab = resize_nearest_legacy(labels, RH, RW)
  if preserve_aspect_ratio:
    oy, ox = offset
    pro = pro[:, oy:oy + th, ox:ox + tw]
    lab = lab[:, oy:oy + th, ox:ox + tw]
  return pro, lab
