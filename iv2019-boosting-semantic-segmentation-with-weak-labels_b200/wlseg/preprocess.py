"""Input-side resize + random crop on the device: code/input_pipelines/utils.py:181-247 `resize_images_and_labels`
(over code/utils/utils.py:540-605 `resize_images_or_labels`, mode 'max', crop None).

  images  fp32 [N, H, W, C]                  tf.image.resize_images(BILINEAR), align_corners=False
  labels  int32 [N, H, W] (class ids) or fp32 [N, H, W, 15] (dense weak labels)   NEAREST_NEIGHBOR
  --preserve_aspect_ratio: both are resized by the SAME factor max(th / H, tw / W) (float64, as the reference's
  implicit cast) to ceil(factor * size) - the tightest size the target fits in - and one random window of the
  target size is cut out of both (offset ~ U{0 .. extra}, tf.random_uniform(maxval=extra + 1) per axis).

Launch sequencing only; the arithmetic is wlseg_resize_crop (csrc/preproc.cu), which reads the crop window straight
out of the source tensor.
"""

import math

import torch

from wlseg import ops


def resized_size(H, W, target_size, preserve_aspect_ratio):
  """utils/utils.py:571-589: the size `resize_images_or_labels` resizes to before any crop."""
  th, tw = int(target_size[0]), int(target_size[1])
  if not preserve_aspect_ratio:
    return th, tw
  factor = max(th / H, tw / W)
  return int(math.ceil(factor * H)), int(math.ceil(factor * W))


def resize_images_and_labels(images, labels, target_size, preserve_aspect_ratio=False, offset=None, generator=None):
  """-> (proimages [N, th, tw, C], prolabels [N, th, tw(, 15)]).  `offset` fixes the crop origin (tests, replay);
  otherwise it is drawn from `generator` (host RNG) exactly as the reference draws it: one integer per axis."""
  assert images.dim() == 4 and labels.dim() in (3, 4), 'images NHWC, labels [N, H, W] or [N, H, W, C]'
  assert tuple(images.shape[1:3]) == tuple(labels.shape[1:3]), 'images and labels must have the same spatial size'
  H, W = int(images.shape[1]), int(images.shape[2])
  th, tw = int(target_size[0]), int(target_size[1])
  RH, RW = resized_size(H, W, (th, tw), preserve_aspect_ratio)
  if preserve_aspect_ratio:
    if offset is None:
      offset = (int(torch.randint(0, RH - th + 1, (), generator=generator)),
                int(torch.randint(0, RW - tw + 1, (), generator=generator)))
  else:
    offset = (0, 0)
  pro = ops.resize_crop(images.to(torch.float32), (RH, RW), offset, (th, tw), 'bilinear')
  lab = ops.resize_crop(labels, (RH, RW), offset, (th, tw), 'nearest')
  return pro, lab
