"""Parity of the fused head-forward and loss forward/backward kernels with the oracle.

Tolerances (floating point, fp32 kernels): 1e-4 relative on logits / probabilities / losses
(north star fp32 check mode); decisions must agree except at numerical near-ties, which the
seeded inputs do not contain (asserted exactly); gradients: cosine >= 0.99999 and 1e-4 relative.
"""

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import metrics as ometrics
from oracle import network as onet
from oracle import tfops
from oracle import weak_labels as oweak
from oracle.tables import TABLES

pytestmark = pytest.mark.gpu


def _hier(dataset):
  from wlseg import hierarchy, problem_defs
  return hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])


def _lowres(hier, N, h, w, seed, scale=3.0):
  g = torch.Generator().manual_seed(seed)
  ct = hier.total_channels
  logits = torch.zeros(N, h, w, hier.logits_pitch)
  logits[..., :ct] = torch.randn(N, h, w, ct, generator=g) * scale
  return logits


def _split(hier, t):
  c1, cv, ch = hier.head_widths
  return t[..., :c1], t[..., c1:c1 + cv], t[..., c1 + cv:c1 + cv + ch]


@pytest.mark.parametrize('dataset,N,h,w,H,W', [
    ('cityscapes', 2, 8, 16, 64, 128),
    ('cityscapes', 1, 7, 9, 50, 70),       # ragged: not a multiple of the tile, odd scale
    ('vistas', 1, 9, 12, 68, 95),
    ('cityscapes', 1, 4, 4, 4, 4),         # identity resize
])
def test_head_fwd_matches_oracle(cuda, dataset, N, h, w, H, W):
  from wlseg import network
  hier = _hier(dataset)
  logits = _lowres(hier, N, h, w, seed=h * 100 + W)
  want_keys = ('decisions', 'l1_decisions', 'l2_vehicle_decisions', 'l2_human_decisions', 'l1_probabilities',
               'l2_vehicle_probabilities', 'l2_human_probabilities', 'logits')
  net = network.Network.__new__(network.Network)
  net.hier, net.hstruct, net.dev = hier, hier.as_struct(), cuda
  got = net.head(logits.to(cuda), H, W, want_keys)
  torch.cuda.synchronize()
  l1, l2v, l2h = [tfops.resize_bilinear(z.contiguous(), H, W) for z in _split(hier, logits)]
  ref = onet.compose_predictions(l1, l2v, l2h, dataset)
  for key in ('l1_logits', 'l2_vehicle_logits', 'l2_human_logits', 'l1_probabilities', 'l2_vehicle_probabilities',
              'l2_human_probabilities'):
    a, b = got[key].cpu(), ref[key]
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max())), key
  for key in ('decisions', 'l1_decisions', 'l2_vehicle_decisions', 'l2_human_decisions'):
    assert torch.equal(got[key].cpu(), ref[key]), key


def test_head_argmax_first_index_on_ties(cuda):
  from wlseg import network
  hier = _hier('cityscapes')
  logits = torch.zeros(1, 2, 2, hier.logits_pitch)  # all-equal logits: every head must answer 0
  net = network.Network.__new__(network.Network)
  net.hier, net.hstruct, net.dev = hier, hier.as_struct(), cuda
  got = net.head(logits.to(cuda), 16, 16, ('decisions', 'l1_decisions', 'l2_vehicle_decisions'))
  assert int(got['l1_decisions'].abs().sum()) == 0 and int(got['l2_vehicle_decisions'].abs().sum()) == 0
  assert int((got['decisions'] != hier.l1_2common[0]).sum()) == 0


def _weak_labels(rng, n, H, W, kind):
  out = np.zeros((n, H, W, 15), dtype=np.float32)
  for i in range(n):
    if kind == 'bbox':
      boxes = []
      for _ in range(int(rng.integers(1, 8))):
        x0, x1 = sorted(rng.random(2))
        y0, y1 = sorted(rng.random(2))
        boxes.append((int(rng.integers(0, 14)), x0, min(x1, 0.999), y0, min(y1, 0.999)))
      out[i] = oweak.bbox_labels(boxes, H, W)
    else:
      cids = list(rng.choice(14, size=int(rng.integers(0, 4)), replace=False))
      out[i] = oweak.image_labels(cids, H, W)
  return torch.from_numpy(out)


@pytest.mark.parametrize('dataset,ns,nb,ni,h,w,H,W', [
    ('cityscapes', 2, 0, 0, 8, 12, 64, 96),
    ('cityscapes', 1, 2, 1, 6, 20, 45, 150),     # mixed strong + bbox + image, ragged tile edges
    ('vistas', 1, 1, 1, 5, 7, 40, 56),
    ('vistas', 2, 0, 0, 17, 30, 136, 239),       # chunked wide-hierarchy kernel: several CTAs, ragged width, 3 strips
    ('cityscapes', 3, 0, 0, 25, 40, 200, 317),   # loss_strong_kernel: several strips per image, ragged tile edges
    ('cityscapes', 0, 1, 1, 4, 6, 32, 48),       # weak only: L1 loss must be exactly 0
])
def test_loss_fwd_bwd_matches_oracle(cuda, dataset, ns, nb, ni, h, w, H, W):
  from wlseg import ops
  hier = _hier(dataset)
  hs = hier.as_struct()
  B = ns + nb + ni
  rng = np.random.default_rng(h * 31 + W)
  logits = _lowres(hier, B, h, w, seed=H + w, scale=2.0)
  ncls = TABLES[dataset]['num_classes']
  strong = torch.from_numpy(rng.integers(0, ncls, size=(ns, H, W), dtype=np.int32))
  # make the vehicle / human classes frequent enough to matter
  strong[:, : H // 3] = int(np.flatnonzero(np.array(hier.pp2veh) != hier.Cv - 1)[0])
  bbox = _weak_labels(rng, nb, H, W, 'bbox')
  image = _weak_labels(rng, ni, H, W, 'image')

  # ---- oracle: upsample -> predictions -> losses, autograd back to the low-res logits
  low = logits[..., :hier.total_channels].clone().requires_grad_(True)
  l1, l2v, l2h = [tfops.resize_bilinear(z, H, W) for z in _split(hier, low)]
  pred = {'l1_logits': l1, 'l2_vehicle_logits': l2v, 'l2_human_logits': l2h,
          'l1_decisions': tfops.argmax_first(tfops.softmax(l1.detach()))}
  labels = {'prolabels_per_pixel': strong, 'prolabels_per_bbox': bbox, 'prolabels_per_image': image}
  if ns == 0:
    # the reference always has strong images; restate the weak-only limit with an empty strong part
    labels['prolabels_per_pixel'] = torch.zeros(0, H, W, dtype=torch.int32)
  ref = olosses.define_losses(pred, labels, dataset)
  ref['segmentation'].backward()

  # ---- kernels
  dl = torch.zeros(B, h, w, hier.logits_pitch, device=cuda)
  sums = torch.zeros(3, dtype=torch.float64, device=cuda)
  counts = torch.zeros(3, dtype=torch.float64, device=cuda)
  out = torch.zeros(4, device=cuda)
  ops.loss_fwd_bwd(hs, logits.to(cuda), H, W, strong.to(cuda) if ns else None, bbox.to(cuda) if nb else None,
                   image.to(cuda) if ni else None, sums, counts, dl)
  ops.loss_finalize(hs, sums, counts, 0.1, 1.0, dl, out)
  torch.cuda.synchronize()
  got = out.cpu()
  want = torch.stack([ref['l1_segmentation'], ref['l2_vehicle_segmentation'], ref['l2_human_segmentation'],
                      ref['segmentation']]).detach()
  cnt = counts.cpu()
  assert cnt[0] == ref['counts']['l1'] and cnt[1] == ref['counts']['l2_vehicle'] and cnt[2] == ref['counts']['l2_human']
  assert torch.allclose(got, want, rtol=1e-4, atol=1e-6), (got, want)
  if ns == 0:
    assert got[0] == 0
  g = dl.cpu()[..., :hier.total_channels]
  assert float(dl.cpu()[..., hier.total_channels:].abs().max() if hier.logits_pitch > hier.total_channels else 0) == 0
  gr = low.grad
  assert float((g - gr).abs().max()) <= 1e-4 * float(gr.abs().max()) + 1e-9
  cos = float((g * gr).sum() / (g.norm() * gr.norm() + 1e-30))
  assert cos >= 0.99999


def test_loss_known_answer_segment_sum(cuda):
  """Worked example of define_losses_hierarchical.py:112-113: a pixel inside one human box and
  one vehicle box gives the vehicle head the target 1/2 vehicle-class + 1/2 void; with the L1
  argmax on 'vehicle' the pixel is supervised by the vehicle head only."""
  from wlseg import ops
  hier = _hier('cityscapes')
  hs = hier.as_struct()
  H = W = 8
  lab = oweak.bbox_labels([(2, 0.0, 0.999, 0.0, 0.999), (6, 0.0, 0.999, 0.0, 0.999)], H, W)  # car + human
  assert np.allclose(lab[0, 0, [2, 6]], 0.5)
  logits = torch.zeros(1, 1, 1, hier.logits_pitch)
  logits[..., hier.cid_l1_vehicle] = 5.0
  dl = torch.zeros_like(logits, device=cuda)
  sums = torch.zeros(3, dtype=torch.float64, device=cuda)
  counts = torch.zeros(3, dtype=torch.float64, device=cuda)
  ops.loss_fwd_bwd(hs, logits.to(cuda), H, W, None, torch.from_numpy(lab)[None].to(cuda), None, sums, counts, dl)
  torch.cuda.synchronize()
  assert counts.cpu().tolist() == [0.0, float(H * W), 0.0]
  # uniform vehicle logits: CE = -(1/2 log(1/7) + 1/2 log(1/7)) = log 7 per pixel
  assert abs(float(sums[1]) / (H * W) - np.log(7.0)) < 1e-5


def test_loss_from_compact_weak_labels_equals_dense_at_training_size(cuda):
  """wlseg_loss_fwd_bwd_lists (weak labels expanded per pixel in the kernel from (class, box) lists and image-level
  class vectors) against the dense path fed with the rasterised / tiled labels, 2 + 4 + 2 images of 256 x 384 with
  up to 12 boxes each: counts identical, losses and the low-res gradient equal to the atomics' rounding (1e-6)."""
  from wlseg import ops, synthetic
  hier = _hier('cityscapes')
  hs = hier.as_struct()
  H, W, h, w = 256, 384, 32, 48
  src = synthetic.SyntheticInputs(20, cuda, seed=5)
  logits = _lowres(hier, 8, h, w, seed=3, scale=2.0).to(cuda)
  strong = src.strong_labels(2, H, W)
  coords, cids = src.bbox_lists(4)
  vec = src.image_vectors(2)
  outs = []
  for compact in (False, True):
    dl = torch.zeros_like(logits)
    sums = torch.zeros(3, dtype=torch.float64, device=cuda)
    counts = torch.zeros(3, dtype=torch.float64, device=cuda)
    out = torch.zeros(4, device=cuda)
    if compact:
      ops.loss_fwd_bwd_lists(hs, logits, H, W, strong, coords, cids, vec, sums, counts, dl)
    else:
      ops.loss_fwd_bwd(hs, logits, H, W, strong, ops.rasterize_bbox_labels(coords, cids, H, W),
                       ops.tile_image_labels(vec, H, W), sums, counts, dl)
    ops.loss_finalize(hs, sums, counts, 0.1, 1.0, dl, out)
    torch.cuda.synchronize()
    outs.append((out.cpu(), counts.cpu(), dl.cpu()))
  (o0, c0, d0), (o1, c1, d1) = outs
  assert torch.equal(c0, c1) and float(c0[1]) > 0 and float(c0[2]) > 0
  assert torch.allclose(o0, o1, rtol=1e-6, atol=1e-7)
  assert float((d0 - d1).abs().max()) <= 1e-6 * float(d0.abs().max())


@pytest.mark.parametrize('dataset,N,h,w,H,W,with_lut', [
    ('cityscapes', 2, 8, 16, 64, 128, False),
    ('cityscapes', 1, 7, 9, 50, 70, True),        # ragged tile, odd scale, decisions remapped through a LUT
    ('cityscapes', 3, 33, 41, 264, 328, False),   # several CTAs per image, runs crossing strip boundaries
    ('vistas', 1, 9, 12, 68, 95, False),
    ('cityscapes', 4, 128, 256, 1024, 2048, True),   # BASELINE configs[1] step
])
def test_head_confmat_equals_oracle_composition_and_histogram(cuda, dataset, N, h, w, H, W, with_lut):
  """wlseg_head_confmat (evaluation tail in one launch) against the ORACLE: decisions from the oracle's
  resize_bilinear + compose_predictions, confusion matrix from the oracle histogram of (labels, lut[decisions]) -
  both bit-exact - and against the two-launch product path (wlseg_head_fwd + wlseg_confmat_accumulate).  Labels carry
  out-of-range ids (counted in `invalid`, skipped) and blocky runs as a segmentation map has them."""
  from wlseg import network, ops
  hier = _hier(dataset)
  C = hier.num_classes
  logits = _lowres(hier, N, h, w, seed=h * 100 + W + 1)
  # make the L2 heads matter: push the vehicle / human super-classes up in a third of the cells
  g = torch.Generator().manual_seed(5)
  cells = torch.rand(N, h, w, generator=g)
  logits[..., hier.as_struct().cid_l1_vehicle] += torch.where(cells < 0.2, 6.0, 0.0)
  logits[..., hier.as_struct().cid_l1_human] += torch.where((cells >= 0.2) & (cells < 0.33), 6.0, 0.0)
  rng = np.random.default_rng(H + W)
  blocks = rng.integers(-1, C + 1, size=(N, -(-H // 16), -(-W // 24)), dtype=np.int32)   # -1 and C are out of range
  labels = torch.from_numpy(np.repeat(np.repeat(blocks, 16, axis=1), 24, axis=2)[:, :H, :W].copy())
  lut = None
  if with_lut:
    lut = torch.from_numpy(rng.integers(0, C, size=C, dtype=np.int32))
    lut[3] = C + 7          # a decision that maps outside the matrix: counted as invalid
  l1, l2v, l2h = [tfops.resize_bilinear(z.contiguous(), H, W) for z in _split(hier, logits)]
  ref_dec = onet.compose_predictions(l1, l2v, l2h, dataset)['decisions']
  mapped = ref_dec if lut is None else lut[ref_dec.long()]
  ok = (labels >= 0) & (labels < C) & (mapped >= 0) & (mapped < C)
  want_cm = ometrics.confusion_matrix(labels[ok].numpy(), mapped[ok].numpy(), C)
  want_bad = int((~ok).sum())

  low = logits.to(cuda)
  lab = labels.to(cuda)
  lut_d = None if lut is None else lut.to(cuda)
  hs = hier.as_struct()
  for keep_decisions in (True, False):
    cm = torch.zeros(C, C, dtype=torch.int64, device=cuda)
    bad = torch.zeros(1, dtype=torch.int64, device=cuda)
    dec = torch.full((N, H, W), -7, dtype=torch.int32, device=cuda) if keep_decisions else None
    ops.head_confmat(hs, low, H, W, lab, C, cm, lut_d, bad, dec)
    ops.head_confmat(hs, low, H, W, lab, C, cm, lut_d, bad, dec)       # accumulates
    torch.cuda.synchronize()
    assert np.array_equal(cm.cpu().numpy(), 2 * want_cm)
    assert int(bad.item()) == 2 * want_bad
    if keep_decisions:
      assert torch.equal(dec.cpu(), ref_dec)
  # decisions only (labels None), and the two-launch path on the same inputs
  dec2 = torch.empty((N, H, W), dtype=torch.int32, device=cuda)
  ops.head_confmat(hs, low, H, W, None, C, None, None, None, dec2)
  net = network.Network.__new__(network.Network)
  net.hier, net.hstruct, net.dev = hier, hs, cuda
  dec3 = net.head(low, H, W, ('decisions',))['decisions']
  cm3 = torch.zeros(C, C, dtype=torch.int64, device=cuda)
  ops.confmat_accumulate(lab, dec3, C, cm3, lut_d)
  torch.cuda.synchronize()
  assert torch.equal(dec2.cpu(), ref_dec) and torch.equal(dec3.cpu(), ref_dec)
  assert np.array_equal(cm3.cpu().numpy(), want_cm)
