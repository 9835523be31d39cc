"""Parity of the confusion-matrix kernel with the oracle: bit-exact (integer work)."""

import numpy as np
import pytest
import torch

from oracle import metrics as ometrics

pytestmark = pytest.mark.gpu


def _run(cuda, labels, decs, C, lut=None):
  from wlseg import ops
  cm = torch.zeros(C, C, dtype=torch.int64, device=cuda)
  invalid = torch.zeros(1, dtype=torch.int64, device=cuda)
  lut_t = None if lut is None else torch.tensor(lut, dtype=torch.int32, device=cuda)
  ops.confmat_accumulate(torch.as_tensor(labels, dtype=torch.int32).to(cuda).contiguous(),
                         torch.as_tensor(decs, dtype=torch.int32).to(cuda).contiguous(), C, cm, lut_t, invalid)
  torch.cuda.synchronize()
  return cm.cpu().numpy(), int(invalid.item())


@pytest.mark.parametrize('C', [20, 66, 3])
@pytest.mark.parametrize('n', [0, 1, 3, 5, 1000, 4 * 257 * 129, 1 << 20])
def test_confmat_matches_oracle(cuda, C, n):
  rng = np.random.default_rng(n + C)
  labels = rng.integers(0, C, size=n, dtype=np.int32)
  decs = rng.integers(0, C, size=n, dtype=np.int32)
  got, bad = _run(cuda, labels, decs, C)
  assert bad == 0
  assert np.array_equal(got, ometrics.confusion_matrix(labels, decs, C))
  if n:
    assert np.array_equal(got, ometrics.confusion_matrix_c(labels, decs, C))
  assert got.sum() == n


def test_confmat_segmentation_like_runs(cuda):
  """Long runs of one class (what a segmentation map looks like) stress the warp aggregation."""
  rng = np.random.default_rng(7)
  lab = np.repeat(np.repeat(rng.integers(0, 20, size=(2, 16, 32), dtype=np.int32), 32, axis=1), 32, axis=2)
  dec = np.repeat(np.repeat(rng.integers(0, 20, size=(2, 32, 64), dtype=np.int32), 16, axis=1), 16, axis=2)
  got, bad = _run(cuda, lab, dec, 20)
  assert bad == 0 and np.array_equal(got, ometrics.confusion_matrix(lab, dec, 20))


def test_confmat_accumulates_and_unaligned(cuda):
  from wlseg import ops
  rng = np.random.default_rng(3)
  C = 20
  cm = torch.zeros(C, C, dtype=torch.int64, device=cuda)
  want = np.zeros((C, C), dtype=np.int64)
  for n in (1001, 4097, 17):
    labels = rng.integers(0, C, size=n + 1, dtype=np.int32)
    decs = rng.integers(0, C, size=n + 1, dtype=np.int32)
    lt = torch.from_numpy(labels).to(cuda)[1:]  # 4-byte aligned only -> scalar path
    dt = torch.from_numpy(decs).to(cuda)[1:]
    ops.confmat_accumulate(lt.contiguous() if False else lt, dt, C, cm)
    want += ometrics.confusion_matrix(labels[1:], decs[1:], C)
  torch.cuda.synchronize()
  assert np.array_equal(cm.cpu().numpy(), want)


def test_confmat_lut_and_invalid(cuda):
  """cid remap worked example of define_estimator_hierarchical.py:494-496 fused as a LUT, and
  out-of-range pairs are skipped and counted."""
  lut = ometrics.replacevoids([-1, 1, 1, 0, -1])
  assert lut == [2, 1, 1, 0, 2]
  rng = np.random.default_rng(11)
  labels = rng.integers(0, 3, size=5000, dtype=np.int32)
  decs = rng.integers(0, 5, size=5000, dtype=np.int32)
  got, bad = _run(cuda, labels, decs, 3, lut)
  assert bad == 0
  assert np.array_equal(got, ometrics.confusion_matrix(labels, ometrics.map_decisions_to_new_cids(decs, [-1, 1, 1, 0, -1]), 3))
  labels[::7] = 99
  decs[::11] = -4
  ok = (labels < 3) & (decs >= 0)
  got, bad = _run(cuda, labels, decs, 3, lut)
  assert bad == int((~ok).sum())
  assert np.array_equal(got, ometrics.confusion_matrix(
      labels[ok], ometrics.map_decisions_to_new_cids(decs[ok], [-1, 1, 1, 0, -1]), 3))


def test_confmat_full_size_checksum(cuda):
  """BASELINE config 2 size (4 x 1024 x 2048): total count and marginals (size-independent properties)."""
  from wlseg import ops
  g = torch.Generator(device='cuda').manual_seed(5)
  n = 4 * 1024 * 2048
  labels = torch.randint(0, 20, (n,), dtype=torch.int32, device=cuda, generator=g)
  decs = torch.randint(0, 20, (n,), dtype=torch.int32, device=cuda, generator=g)
  cm = torch.zeros(20, 20, dtype=torch.int64, device=cuda)
  ops.confmat_accumulate(labels, decs, 20, cm)
  torch.cuda.synchronize()
  assert int(cm.sum()) == n
  assert torch.equal(cm.sum(1), torch.bincount(labels.long(), minlength=20))
  assert torch.equal(cm.sum(0), torch.bincount(decs.long(), minlength=20))
  assert torch.equal(cm.reshape(-1), torch.bincount(labels.long() * 20 + decs.long(), minlength=400))
