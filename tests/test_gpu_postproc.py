"""Prediction post-processing kernels (`_resize_predictions`, `_replace_voids`;
code/estimator/define_estimator_hierarchical.py:530-630) against the oracle.

Bar: decisions (integer gather) bit-exact; probabilities 1e-6 absolute (same fp32 lerp order as the oracle's
restatement of the TF-1.12 kernel, up to fma contraction); void replacement bit-exact.
"""

import numpy as np
import pytest
import torch

from oracle import metrics as ometrics
from oracle import tfops
from oracle.tables import TABLES

pytestmark = pytest.mark.gpu

SIZES = [((16, 32), (128, 256)), ((64, 128), (50, 77)), ((33, 47), (33, 47)), ((24, 40), (1, 1)), ((7, 5), (61, 90)),
         ((1, 1), (9, 13)), ((128, 256), (1024, 2048))]


@pytest.mark.parametrize('src,dst', SIZES)
def test_resize_decisions_bit_exact(cuda, src, dst):
  from wlseg import ops
  g = torch.Generator().manual_seed(src[0] * 131 + dst[1])
  decs = torch.randint(0, 66, (2, src[0], src[1]), generator=g, dtype=torch.int32)
  got = ops.resize_decisions(decs.to(cuda), dst[0], dst[1]).cpu()
  want = tfops.resize_nearest(decs, dst[0], dst[1], align_corners=True)
  assert got.dtype == torch.int32 and tuple(got.shape) == (2, dst[0], dst[1])
  assert torch.equal(got, want)


@pytest.mark.parametrize('C', [14, 7, 3, 53])
@pytest.mark.parametrize('src,dst', SIZES[:6])
def test_resize_probabilities(cuda, src, dst, C):
  from wlseg import ops
  g = torch.Generator().manual_seed(src[1] * 17 + dst[0] + C)
  probs = torch.softmax(torch.randn(2, src[0], src[1], C, generator=g) * 3, -1)
  got = ops.resize_probabilities(probs.to(cuda), dst[0], dst[1]).cpu()
  want = tfops.resize_bilinear(probs, dst[0], dst[1], align_corners=True)
  assert tuple(got.shape) == tuple(want.shape)
  assert float((got - want).abs().max()) <= 1e-6


@pytest.mark.parametrize('dataset', ['cityscapes', 'vistas'])
def test_replace_voids_bit_exact(cuda, dataset):
  from wlseg import hierarchy, ops, problem_defs
  hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
  t = TABLES[dataset]
  C1, Cv, Ch = t['head_widths']
  g = torch.Generator().manual_seed(C1)
  N, H, W = 2, 37, 53
  lg = [torch.randn(N, H, W, c, generator=g) for c in (C1, Cv, Ch)]
  for x in lg:
    x[..., -1] += 1.5          # plenty of void winners in every head
  lg[0][..., t['cid_l1_vehicle']] += 1.0
  lg[0][..., t['cid_l1_human']] += 1.0
  p1, pv, ph = (torch.softmax(x, -1) for x in lg)
  d1, dv, dh = (tfops.argmax_first(p).numpy() for p in (p1, pv, ph))
  l1c, vc, hc = (np.asarray(t[k]) for k in ('l1_cids2common_cids', 'l2_vehicle_cids2common_cids', 'l2_human_cids2common_cids'))
  decs = np.where(d1 == t['cid_l1_vehicle'], vc[dv], np.where(d1 == t['cid_l1_human'], hc[dh], l1c[d1])).astype(np.int32)
  void = t['num_classes'] - 1
  assert hier.void_cid == void and 0.2 < float((decs == void).mean()) < 0.9
  want = ometrics.replace_voids_hierarchical(p1.numpy(), pv.numpy(), ph.numpy(), decs, t)
  got = torch.from_numpy(decs).to(cuda)
  ops.replace_voids(hier.as_struct(), p1.to(cuda), pv.to(cuda), ph.to(cuda), got, void)
  got = got.cpu().numpy()
  assert np.array_equal(got, want)
  assert not (got == void).any()                       # no void decision survives ...
  assert np.array_equal(got[decs != void], decs[decs != void])   # ... and nothing else changed


def test_predict_resizes_to_system_size_and_replaces_voids(cuda):
  """PREDICT branch through the Estimator: outputs at (height_system, width_system), equal to the oracle's
  two-step pipeline (network-size predictions -> _resize_predictions -> void replacement)."""
  import argparse
  from oracle import network as onet
  from wlseg import estimator as est, hierarchy, problem_defs
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  s = argparse.Namespace(dtype='fp32', stride_feature_extractor=8, psp_module=False, height_system=90, width_system=150,
                         replace_voids=True, batch_norm_decay=0.9)
  e = est.Estimator(s, hier, device=cuda)
  tf_params = onet.init_params('cityscapes', seed=5, randomize_bn=True, tame=True)
  e.params.load_tf_dict(tf_params)
  from wlseg import network
  e.net = network.Network(e.params, dtype=torch.float32)
  g = torch.Generator().manual_seed(2)
  images = torch.rand(2, 64, 96, 3, generator=g) * 2 - 1
  keys = ['l1_probabilities', 'l2_vehicle_probabilities', 'l2_human_probabilities', 'decisions', 'rawimagespaths']
  outs = list(e.predict([({'proimages': images.to(cuda), 'rawimagespaths': [b'a', b'b']}, None)], keys))
  assert len(outs) == 2 and outs[0]['rawimagespaths'] == b'a'
  ref = onet.Net(tf_params, 'cityscapes').forward(images)
  t = TABLES['cityscapes']
  probs = [tfops.resize_bilinear(ref[k], 90, 150, align_corners=True) for k in keys[:3]]
  decs = tfops.resize_nearest(ref['decisions'], 90, 150, align_corners=True)
  want = ometrics.replace_voids_hierarchical(*(p.numpy() for p in probs), decs.numpy(), t)
  for i in range(2):
    assert outs[i]['decisions'].shape == (90, 150) and outs[i]['decisions'].dtype == np.int32
    for k, p in zip(keys[:3], probs):
      assert outs[i][k].shape == tuple(p.shape[1:])
      assert float(np.abs(outs[i][k] - p[i].numpy()).max()) <= 1e-4
    dis = float((outs[i]['decisions'] != want[i]).mean())
    assert dis <= 2e-3, dis    # fp32 check mode: near-ties only
    assert not (outs[i]['decisions'] == 19).any()
