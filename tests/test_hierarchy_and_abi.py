"""Host-side checks that need no GPU: derived hierarchy tables vs the reference's literals,
problem-definition JSON, and that the C-ABI library loads and exports every declared symbol."""

import ctypes
import os
import re

import pytest

from oracle.tables import TABLES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('dataset', ['cityscapes', 'vistas'])
def test_derived_hierarchy_equals_reference_literals(dataset):
  from wlseg import hierarchy, problem_defs
  pd = problem_defs.GENERATORS[dataset]()
  h = hierarchy.Hierarchy(dataset, pd['cids2labels'])
  t = TABLES[dataset]
  assert h.pp2l1 == t['per_pixel_cids2l1_cids']
  assert h.pp2veh == t['per_pixel_cids2vehicle_cids']
  assert h.pp2hum == t['per_pixel_cids2human_cids']
  assert h.bb2veh == t['per_bbox_cids2vehicle_cids']
  assert h.bb2hum == t['per_bbox_cids2human_cids']
  assert h.bb2l1 == t['per_bbox_cids2l1_cids']
  assert h.l1_2common == t['l1_cids2common_cids']
  assert h.veh2common == t['l2_vehicle_cids2common_cids']
  assert h.hum2common == t['l2_human_cids2common_cids']
  assert h.head_widths == t['head_widths']
  assert (h.cid_l1_vehicle, h.cid_l1_human) == (t['cid_l1_vehicle'], t['cid_l1_human'])
  s = h.as_struct()
  assert (s.C1, s.Cv, s.Ch) == t['head_widths'] and s.num_classes == t['num_classes']
  assert list(s.pp_to_l1)[:t['num_classes']] == t['per_pixel_cids2l1_cids']
  assert h.logits_pitch % 8 == 0 and h.logits_pitch >= h.total_channels


def test_problem_definitions_schema_and_counts():
  from wlseg import problem_defs
  c = problem_defs.cityscapes()
  assert len(c['lids2cids']) == 34 and max(c['lids2cids']) == 18 and c['lids2cids'].count(-1) == 15
  assert c['cids2labels'][-1] == 'void' and len(c['cids2labels']) == 20 == len(c['cids2colors']) == len(c['cids2lids'])
  assert c['cids2lids'][:3] == [7, 8, 11]
  v = problem_defs.vistas()
  assert len(v['lids2cids']) == 66 and v['lids2cids'][-1] == -1 and len(v['cids2labels']) == 66
  for path in problem_defs.write_all():
    assert problem_defs.load(path)['version'] == 2.0


def test_library_loads_and_exports_every_declared_symbol():
  """No compute calls (no GPU here): dlopen + symbol table vs include/wlseg.h."""
  from wlseg import ops
  header = open(os.path.join(ROOT, 'include', 'wlseg.h')).read()
  declared = set(re.findall(r'\b(wlseg_[a-z0-9_]+)\s*\(', header))
  declared -= {'wlseg_conv_params', 'wlseg_hierarchy', 'wlseg_stream_t'}
  assert declared == set(ops.EXPORTED_SYMBOLS), declared ^ set(ops.EXPORTED_SYMBOLS)
  lib = ops.lib()
  for name in declared:
    assert getattr(lib, name) is not None
  assert lib.wlseg_version() == 100
  assert ctypes.sizeof(ops.Hierarchy) == 4 * (5 + 64 + 16 + 8 + 1 + 240 + 30)
  assert ctypes.sizeof(ops.ConvParams) == 4 * 25   # 24 fields of round-1 start + `reverse`
  assert ctypes.sizeof(ops.BnFinalizeArgs) == 8 + 4 * 4 + 9 * 8   # wlseg_bn_finalize_args: count, 3 floats + pad, 9 pointers


def test_invalid_arguments_are_reported_not_crashed():
  """Argument validation happens before any CUDA call, so it is testable without a GPU."""
  from wlseg import ops
  lib = ops.lib()
  rc = lib.wlseg_confmat_accumulate(None, None, 10, 0, None, 0, None, None, None)
  assert rc < 0 and b'cm is NULL' in lib.wlseg_last_error()
  rc = lib.wlseg_confmat_accumulate(None, None, 10, 500, None, 0, 1, None, None)
  assert rc < 0 and b'num_classes' in lib.wlseg_last_error()
  p = ops.ConvParams()
  rc = lib.wlseg_conv2d_fprop(ctypes.byref(p), None, None, None, None, None, None, None, None, None)
  assert rc < 0 and b'bad shape' in lib.wlseg_last_error()
  assert lib.wlseg_conv2d_tcgen05_supported(ctypes.byref(p)) in (0, 1)


def test_product_does_not_import_the_oracle():
  pkg = os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200', 'wlseg')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith('.py'):
        src = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f'{f} imports the oracle'


def test_predict_exports_palette_pngs(tmp_path):
  """code/predict.py:137-164: label-id, colour and overlapped PNG exports are palette look-ups on the decisions."""
  import argparse
  import numpy as np
  from PIL import Image
  from wlseg import cli, problem_defs
  pd = problem_defs.cityscapes()
  s = argparse.Namespace(results_dir=str(tmp_path), export_lids_images=True, export_color_decisions=True,
                         export_overlapped_color_decisions=True, inference_problem_def=pd)
  rng = np.random.default_rng(0)
  decs = rng.integers(0, 20, (6, 9), dtype=np.int32)
  raw = rng.integers(0, 256, (6, 9, 3), dtype=np.uint8)
  ids = np.array(pd['cids2lids'], dtype=np.uint8)
  col = np.array(pd['cids2colors'], dtype=np.uint8)
  cli.export_outputs({'decisions': decs, 'rawimages': raw, 'rawimagespaths': b'/data/frankfurt_000001.png'}, s, ids, col)
  assert np.array_equal(np.asarray(Image.open(tmp_path / 'frankfurt_000001_result_lids.png')), ids[decs])
  assert np.array_equal(np.asarray(Image.open(tmp_path / 'frankfurt_000001_result_color.png')), col[decs])
  want = (0.5 * raw + 0.5 * col[decs]).astype(np.uint8)
  assert np.array_equal(np.asarray(Image.open(tmp_path / 'frankfurt_000001_result_overlapped_color.png')), want)


def test_real_image_predict_input(tmp_path):
  """dataset_agnostic_predict_input.py:88-154: recursive image discovery, RGB conversion, [-1, 1) scaling, legacy
  align_corners=False bilinear resize to the feature-extractor size, one image per batch with raw image and path."""
  import argparse
  import numpy as np
  import torch
  from PIL import Image
  from wlseg import image_input
  rng = np.random.default_rng(1)
  (tmp_path / 'sub').mkdir()
  a = rng.integers(0, 256, (6, 8, 3), dtype=np.uint8)
  Image.fromarray(a).save(tmp_path / 'a.png')
  Image.fromarray(rng.integers(0, 256, (5, 5), dtype=np.uint8), mode='L').save(tmp_path / 'sub' / 'grey.png')
  (tmp_path / 'notes.txt').write_text('not an image')
  params = argparse.Namespace(predict_dir=str(tmp_path), height_feature_extractor=12, width_feature_extractor=16, Nb=1)
  batches = list(image_input.predict_input_fn(None, params))
  assert len(batches) == 2
  by_name = {os.path.basename(f['rawimagespaths'][0].decode()): f for f, _ in batches}
  f = by_name['a.png']
  assert tuple(f['proimages'].shape) == (1, 12, 16, 3) and f['proimages'].dtype == torch.float32
  assert tuple(f['rawimages'].shape) == (1, 6, 8, 3) and np.array_equal(f['rawimages'][0].numpy(), a)
  assert float(f['proimages'].min()) >= -1.0 and float(f['proimages'].max()) <= 1.0
  # exact 2x upscale, legacy mapping: output (2i, 2j) is input (i, j), odd positions are midpoints, the last one is
  # clamped to the border
  x = a.astype(np.float32) / 255.0
  want00 = (x[0, 0] - 0.5) / 0.5
  assert np.allclose(f['proimages'][0, 0, 0].numpy(), want00, atol=1e-6)
  assert np.allclose(f['proimages'][0, 2, 4].numpy(), (x[1, 2] - 0.5) / 0.5, atol=1e-6)
  assert np.allclose(f['proimages'][0, 1, 0].numpy(), ((x[0, 0] + x[1, 0]) / 2 - 0.5) / 0.5, atol=1e-6)
  assert np.allclose(f['proimages'][0, 11, 15].numpy(), (x[5, 7] - 0.5) / 0.5, atol=1e-6)
  assert tuple(by_name['grey.png']['rawimages'].shape) == (1, 5, 5, 3)          # converted to RGB
  # identity size: no resampling at all
  params2 = argparse.Namespace(predict_dir=str(tmp_path), height_feature_extractor=6, width_feature_extractor=8, Nb=1, steps=1)
  only = list(image_input.predict_input_fn(None, params2))
  assert len(only) == 1
