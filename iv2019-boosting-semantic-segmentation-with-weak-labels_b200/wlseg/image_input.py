"""Real-image input for predict.py: the dataset-agnostic pipeline of
code/input_pipelines/dataset_agnostic/dataset_agnostic_predict_input.py:88-154 without tf.data.

  _predict_image_generator (:88-107)  every png / jpg / jpeg / ppm under predict_dir (recursive), converted to RGB
  _predict_preprocess (:109-119)      uint8 -> float32 in [0, 1] -> bilinear resize to (height_feature_extractor,
                                      width_feature_extractor) -> [-1, 1)  (from_0_1_to_m1_1, input_pipelines/utils.py:95-110)
  predict_input (:121-154)            one image per batch; features = {rawimages, proimages, rawimagespaths}

Input plumbing on the host (PIL + torch CPU), like wlseg/synthetic.py it only honours the tensor contract of the hot
path; the network, the prediction resize back to the raw image size and the exports run on the device.
[TF-1.12] tf.image.resize_images(bilinear) defaults to align_corners=False with the legacy mapping src = dst * in/out
(no half-pixel centres), lerp in fp32.
"""

import glob
import os

import numpy as np
import torch

SUPPORTED_EXTENSIONS = ('png', 'PNG', 'jpg', 'JPG', 'jpeg', 'JPEG', 'ppm', 'PPM')


def list_images(predict_dir):
  fnames = []
  for se in SUPPORTED_EXTENSIONS:
    fnames.extend(glob.glob(os.path.join(predict_dir, '**', '*.' + se), recursive=True))
  return fnames


def resize_bilinear_legacy(img, out_h, out_w):
  """[TF-1.12] ResizeBilinear, align_corners=False: src = dst * (in / out); lo = floor(src), hi = min(lo + 1, in - 1).
  img: float32 [H, W, C] (host) -> [out_h, out_w, C]."""
  h, w = img.shape[0], img.shape[1]
  if (h, w) == (out_h, out_w):
    return img
  ys = torch.arange(out_h, dtype=torch.float32) * torch.tensor(h / out_h, dtype=torch.float32)
  xs = torch.arange(out_w, dtype=torch.float32) * torch.tensor(w / out_w, dtype=torch.float32)
  y0, x0 = torch.floor(ys).long(), torch.floor(xs).long()
  y1, x1 = torch.clamp(y0 + 1, max=h - 1), torch.clamp(x0 + 1, max=w - 1)
  ty, tx = (ys - y0.float()).view(out_h, 1, 1), (xs - x0.float()).view(1, out_w, 1)
  top = img[y0][:, x0] + (img[y0][:, x1] - img[y0][:, x0]) * tx
  bot = img[y1][:, x0] + (img[y1][:, x1] - img[y1][:, x0]) * tx
  return top + (bot - top) * ty


def predict_input_fn(config, params):
  """Drop-in `input_fn(config, params)` for SemanticSegmentation({'predict': ...}) over the images of params.predict_dir."""
  del config
  from PIL import Image
  if getattr(params, 'preserve_aspect_ratio', False):
    raise NotImplementedError('prediction with --preserve_aspect_ratio is not implemented.')
  fnames = list_images(params.predict_dir)
  print(f"Found {len(fnames)} images.")
  if getattr(params, 'Nb', 1) > 1:
    print("\n\nBatching for inference is disabled (in case input images don't have the same size).\n\n")
  hf, wf = params.height_feature_extractor, params.width_feature_extractor
  limit = getattr(params, 'steps', None)
  for i, fname in enumerate(fnames):
    if limit and i >= limit:
      break
    im = Image.open(fname)
    if im.mode != 'RGB':
      print(f"{fname} [{im}] didn't comply with specs. Trying to transform it, otherwise it will ignore it.")
      if im.mode not in ('L', 'P', 'RGBA'):
        continue
      im = im.convert(mode='RGB')
    raw = torch.from_numpy(np.asarray(im, dtype=np.uint8).copy())                 # [H, W, 3]
    # tf.image.convert_image_dtype [TF-1.12]: cast, then MULTIPLY by the float32 constant 1 / 255 (not a division)
    img = raw.to(torch.float32) * torch.tensor(1.0 / 255.0, dtype=torch.float32)
    pro = (resize_bilinear_legacy(img, hf, wf) - 0.5) / 0.5                       # from_0_1_to_m1_1
    yield {'proimages': pro.unsqueeze(0).contiguous(), 'rawimages': raw.unsqueeze(0),
           'rawimagespaths': [fname.encode('utf-8')]}, None
