"""Golden vectors of the dataset-agnostic predict input produced by running the reference's own functions:
tests/golden/reference_predict_input_run.npz.

    python tests/golden/make_reference_predict_input_fixtures.py       # needs /root/reference; run in the build container

`code/input_pipelines/dataset_agnostic/dataset_agnostic_predict_input.py::_predict_image_generator` (:88-107, pure
Python + PIL) and `_predict_preprocess` (:109-119, over utils.utils.resize_images_or_labels and
input_pipelines.utils.from_0_1_to_m1_1) are called UNMODIFIED, with tests/golden/tf_shim first on sys.path, on a small
directory tree written by `write_images` (RGB / grey / palette / RGBA PNGs, an upper-case extension, a PPM, a nested
directory, a non-image file).  Restated TF calls: tf.image.convert_image_dtype (uint8 -> float32: cast, then multiply by
the float32 constant 1 / 255) and the legacy `align_corners=False` bilinear resize of the earlier fixtures.
Stored per image (keyed by its path relative to the directory): the raw RGB array the generator yields and the processed
image.  tests/test_reference_fixtures.py runs wlseg.image_input.predict_input_fn on the same tree.
"""

import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('WLSEG_REFERENCE', '/root/reference/code')
OUT = os.path.join(HERE, 'reference_predict_input_run.npz')
SEED, HF, WF = 61, 24, 40


def write_images(root):
  """The directory the reference and the product both read (lossless formats only, so that the tree is reproducible)."""
  from PIL import Image
  rng = np.random.default_rng(SEED)
  os.makedirs(os.path.join(root, 'nested', 'deeper'))
  Image.fromarray(rng.integers(0, 256, (30, 50, 3), dtype=np.uint8)).save(os.path.join(root, 'rgb.png'))
  Image.fromarray(rng.integers(0, 256, (24, 40, 3), dtype=np.uint8)).save(os.path.join(root, 'same_size.PNG'))
  Image.fromarray(rng.integers(0, 256, (17, 23), dtype=np.uint8), mode='L').save(os.path.join(root, 'nested', 'grey.png'))
  Image.fromarray(rng.integers(0, 256, (40, 31, 4), dtype=np.uint8), mode='RGBA').save(os.path.join(root, 'nested', 'deeper', 'rgba.png'))
  pal = Image.fromarray(rng.integers(0, 8, (21, 33), dtype=np.uint8), mode='P')
  pal.putpalette([int(v) for v in rng.integers(0, 256, 24)])
  pal.save(os.path.join(root, 'nested', 'palette.png'))
  Image.fromarray(rng.integers(0, 256, (48, 80, 3), dtype=np.uint8)).save(os.path.join(root, 'nested', 'double.ppm'))
  with open(os.path.join(root, 'notes.txt'), 'w') as fp:
    fp.write('not an image\n')


def main():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import torch
  import tensorflow as tf
  assert tf.__version__.endswith('shim')

  def convert_image_dtype(image, dtype, saturate=False, name=None):
    """[TF-1.12] integer -> float: math_ops.multiply(cast(image, dtype), 1. / image.dtype.max)."""
    assert dtype == tf.float32 and image.dtype == torch.uint8
    return tf.as_tf(image.to(torch.float32) * torch.tensor(1.0 / 255.0, dtype=torch.float32))
  tf.image.convert_image_dtype = convert_image_dtype
  from input_pipelines.dataset_agnostic import dataset_agnostic_predict_input as dp
  params = types.SimpleNamespace(height_feature_extractor=HF, width_feature_extractor=WF, preserve_aspect_ratio=False)
  out = {'size': np.asarray([HF, WF], dtype=np.int32)}
  with tempfile.TemporaryDirectory() as root:
    write_images(root)
    params.predict_dir = root
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
      items = list(dp._predict_image_generator(params))
    rel = []
    for im, path, height, width in items:
      raw = np.asarray(im, dtype=np.uint8)             # what tf.data's from_generator converts the PIL image to
      assert raw.shape == (height, width, 3)
      pro = dp._predict_preprocess(tf.as_tf(torch.from_numpy(raw.copy())), params)
      key = os.path.relpath(path.decode('utf-8'), root)
      rel.append(key)
      out[f'raw/{key}'] = raw
      out[f'pro/{key}'] = torch.Tensor(pro).numpy().astype(np.float32)
    out['paths'] = np.asarray('\n'.join(sorted(rel)))
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes;', sorted(rel))


if __name__ == '__main__':
  main()
