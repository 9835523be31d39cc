"""Parity of the bandwidth kernels (max-pool, batch-norm train fwd/bwd, SGD-momentum, packing)
with the oracle's TF-1.12 restatements.  fp32 storage: 1e-5 relative; bf16 storage: one bf16 ulp
of the fp32 result (values are compared after rounding the oracle result to bf16)."""

import numpy as np
import pytest
import torch

from oracle import optimizer as oopt
from oracle import tfops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape,k,s', [((2, 12, 16, 64), 3, 2), ((1, 9, 7, 8), 3, 2), ((1, 8, 8, 16), 1, 2), ((2, 37, 45, 64), 3, 2)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_maxpool_same_fwd_bwd(cuda, shape, k, s, dtype):
  from wlseg import ops
  g = torch.Generator().manual_seed(1)
  x = torch.randn(shape, generator=g).to(dtype).float()
  x[0, :4, :4] = 0.0  # ties (post-ReLU zeros): gradient must go to the first maximum
  xr = x.clone().requires_grad_(True)
  yr = tfops.max_pool_same(xr, k, s)
  N, P, Q, C = yr.shape
  y = torch.empty((N, P, Q, C), dtype=dtype, device=cuda)
  xd = x.to(dtype).to(cuda)
  ops.maxpool_same_fwd(xd, y, k, s)
  assert torch.equal(y.float().cpu(), yr.detach())
  dy = torch.randn(yr.shape, generator=g).to(dtype).float()
  yr.backward(dy)
  dx = torch.empty_like(xd)
  ops.maxpool_same_bwd(xd, dy.to(dtype).to(cuda), dx, k, s)
  torch.cuda.synchronize()
  want = xr.grad
  tol = 0 if dtype == torch.float32 else 2e-2
  assert float((dx.float().cpu() - want).abs().max()) <= tol * float(want.abs().max()) + 1e-6
  # same result through the forward's argmax map (the training path: no re-read of x in the backward)
  amax = torch.full((N, P, Q, C), 77, dtype=torch.uint8, device=cuda)
  y2 = torch.empty_like(y)
  ops.maxpool_same_fwd(xd, y2, k, s, argmax=amax)
  dx2 = torch.full_like(dx, float('nan'))
  ops.maxpool_same_bwd(None, dy.to(dtype).to(cuda), dx2, k, s, argmax=amax)
  torch.cuda.synchronize()
  assert torch.equal(y2, y) and int(amax.max()) < k * k
  assert torch.equal(dx2, dx)


@pytest.mark.parametrize('C', [64, 24, 256])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_batch_norm_train_fwd_bwd(cuda, C, dtype):
  from wlseg import ops
  g = torch.Generator().manual_seed(C)
  N, H, W = 2, 9, 11
  count = N * H * W
  z = (torch.randn(N, H, W, C, generator=g) * 2 + 0.5).to(dtype).float()
  res = torch.randn(N, H, W, C, generator=g).to(dtype).float()
  gamma = 0.5 + torch.rand(C, generator=g)
  beta = torch.randn(C, generator=g) * 0.1
  mm = torch.randn(C, generator=g) * 0.1
  mv = 0.5 + torch.rand(C, generator=g)
  zr = z.clone().requires_grad_(True)
  rr = res.clone().requires_grad_(True)
  gr = gamma.clone().requires_grad_(True)
  br = beta.clone().requires_grad_(True)
  y_bn, new_mm, new_mv, mean, var = tfops.batch_norm(zr, gr, br, mm, mv, True, decay=0.9, eps=1e-5)
  yr = torch.relu(y_bn + rr)
  dy = torch.randn(N, H, W, C, generator=g).to(dtype).float()
  yr.backward(dy)

  dev = cuda
  zd, resd = z.to(dtype).to(dev), res.to(dtype).to(dev)
  s1 = torch.zeros(C, dtype=torch.float64, device=dev)
  s2 = torch.zeros(C, dtype=torch.float64, device=dev)
  ops.bn_stats(zd, count, C, C, s1, s2)
  gd, bd, mmd, mvd = gamma.to(dev), beta.to(dev), mm.to(dev), mv.to(dev)
  scale, shift = torch.empty(C, device=dev), torch.empty(C, device=dev)
  smean, sinv = torch.empty(C, device=dev), torch.empty(C, device=dev)
  ops.bn_finalize(s1, s2, count, C, gd, bd, 1e-5, 0.9, mmd, mvd, scale, shift, smean, sinv)
  y = torch.empty_like(zd)
  ops.bn_apply(zd, scale, shift, resd, y, count, C, True)
  torch.cuda.synchronize()
  if C % 8 == 0:
    # the fused finalize + apply launch of the training path: identical results, bit for bit
    mm2, mv2 = mm.to(dev), mv.to(dev)
    sc2, sh2, me2, in2 = (torch.empty(C, device=dev) for _ in range(4))
    y2 = torch.empty_like(zd)
    ops.bn_finalize_apply(s1, s2, count, C, gd, bd, 1e-5, 0.9, mm2, mv2, sc2, sh2, me2, in2, zd, resd, y2, True)
    torch.cuda.synchronize()
    assert torch.equal(y2, y) and torch.equal(sc2, scale) and torch.equal(sh2, shift)
    assert torch.equal(me2, smean) and torch.equal(in2, sinv) and torch.equal(mm2, mmd) and torch.equal(mv2, mvd)
  assert torch.allclose(smean.cpu(), mean.detach(), rtol=1e-5, atol=1e-6)
  assert torch.allclose(sinv.cpu(), torch.rsqrt(var.detach() + 1e-5), rtol=1e-5)
  assert torch.allclose(mmd.cpu(), new_mm, rtol=1e-5, atol=1e-6)
  assert torch.allclose(mvd.cpu(), new_mv, rtol=1e-5, atol=1e-6)
  tol = 1e-5 if dtype == torch.float32 else 1e-2
  assert float((y.float().cpu() - yr.detach()).abs().max()) <= tol * float(yr.abs().max())

  dgm = torch.zeros(C, dtype=torch.float64, device=dev)
  dbt = torch.zeros(C, dtype=torch.float64, device=dev)
  dyd = dy.to(dtype).to(dev)
  # the mask uses the oracle's activation so that bf16 rounding of y cannot flip it
  yact = yr.detach().to(dtype).to(dev)
  ops.bn_bwd_reduce(dyd, yact, zd, smean, sinv, count, C, True, dgm, dbt)
  dz, dres = torch.empty_like(zd), torch.empty_like(zd)
  ops.bn_bwd_apply(dyd, yact, zd, smean, sinv, gd, dgm, dbt, count, C, True, dz, dres)
  torch.cuda.synchronize()
  assert torch.allclose(dgm.float().cpu(), gr.grad, rtol=1e-4, atol=1e-4)
  assert torch.allclose(dbt.float().cpu(), br.grad, rtol=1e-4, atol=1e-4)
  assert float((dz.float().cpu() - zr.grad).abs().max()) <= tol * float(zr.grad.abs().max()) + 1e-6
  assert float((dres.float().cpu() - rr.grad).abs().max()) <= tol * float(rr.grad.abs().max()) + 1e-6

  if C % 16 == 0:
    # training path: ReLU layer WITHOUT residual, mask recomputed from z (y not read), processed in
    # two channel slices through `pitch` - must equal the one-shot y-masked result bit for bit
    y2 = torch.empty_like(zd)
    ops.bn_apply(zd, scale, shift, None, y2, count, C, True)
    ref_dg = torch.zeros(C, dtype=torch.float64, device=dev)
    ref_db = torch.zeros(C, dtype=torch.float64, device=dev)
    ops.bn_bwd_reduce(dyd, y2, zd, smean, sinv, count, C, True, ref_dg, ref_db)
    ref_dz = torch.empty_like(zd)
    ops.bn_bwd_apply(dyd, y2, zd, smean, sinv, gd, ref_dg, ref_db, count, C, True, ref_dz)
    dg2 = torch.zeros(C, dtype=torch.float64, device=dev)
    db2 = torch.zeros(C, dtype=torch.float64, device=dev)
    dz2 = torch.full_like(zd, float('nan'))
    h = C // 2
    for sl in (slice(0, h), slice(h, C)):
      ops.bn_bwd_reduce(dyd[..., sl], None, zd[..., sl], smean[sl], sinv[sl], count, h, True, dg2[sl], db2[sl],
                        scale=scale[sl], shift=shift[sl], pitch=C)
      ops.bn_bwd_apply(dyd[..., sl], None, zd[..., sl], smean[sl], sinv[sl], gd[sl], dg2[sl], db2[sl], count, h, True,
                       dz2[..., sl], scale=scale[sl], shift=shift[sl], pitch=C)
    torch.cuda.synchronize()
    # per-thread partial sums are fp32 and the row partition differs between the two launches
    assert torch.allclose(dg2, ref_dg, rtol=1e-4, atol=1e-4) and torch.allclose(db2, ref_db, rtol=1e-4, atol=1e-4)
    assert float((dz2.float() - ref_dz.float()).abs().max()) <= 1e-2 * float(ref_dz.float().abs().max())


@pytest.mark.parametrize('nesterov', [False, True])
def test_sgdm_matches_momentum_optimizer(cuda, nesterov):
  from wlseg import ops
  g = torch.Generator().manual_seed(9)
  n, n_decay = 1003, 640
  w = torch.randn(n, generator=g)
  acc = torch.randn(n, generator=g) * 0.1
  grad = torch.randn(n, generator=g)
  wd, lr, mom = 1.7e-4, 0.01, 0.9
  gfull = grad.clone()
  gfull[:n_decay] += wd * w[:n_decay]
  w_ref, acc_ref = oopt.momentum_step(w, gfull, acc, lr, mom, nesterov)
  reg_ref = 0.5 * wd * float((w[:n_decay].double() ** 2).sum())
  wdv, accd, gd = w.clone().to(cuda), acc.clone().to(cuda), grad.to(cuda)
  wb = torch.zeros(n, dtype=torch.bfloat16, device=cuda)
  lr_dev = torch.tensor([lr], device=cuda)
  reg = torch.zeros(1, dtype=torch.float64, device=cuda)
  ops.sgdm_step(wdv, gd, accd, wb, n_decay, lr_dev, mom, nesterov, wd, 1.0, reg)
  torch.cuda.synchronize()
  assert torch.allclose(wdv.cpu(), w_ref, rtol=1e-6, atol=1e-7)
  assert torch.allclose(accd.cpu(), acc_ref, rtol=1e-6, atol=1e-7)
  assert torch.equal(wb.cpu(), wdv.cpu().to(torch.bfloat16))
  assert abs(float(reg) - reg_ref) <= 1e-6 * reg_ref


def test_conv1_pack_is_exact_rewrite(cuda):
  """The packed R=4,S=1,C=64 convolution equals the 7x7 stride-2 `conv2d_same` (fp64 check on CPU
  of the index map: pack on the GPU, convolve the packed tensor with the rewritten kernel)."""
  from wlseg import hierarchy, network, ops, problem_defs
  g = torch.Generator().manual_seed(2)
  N, H, W = 1, 20, 26
  img = torch.rand(N, H, W, 3, generator=g) * 2 - 1
  img = img.to(torch.bfloat16).float()
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  params = network.Params(hier, cuda)
  params.init_random(3)
  packed = torch.empty((N, H // 2, W // 2, 64), dtype=torch.bfloat16, device=cuda)
  ops.conv1_pack(img.to(cuda), packed)
  w2 = params.conv1_packed_weights(torch.float32).cpu()          # [64, 4, 1, 64]
  w = params.w32('feature_extractor/base/resnet_v1_50/conv1').cpu()  # [64, 7, 7, 3] KRSC
  ref = tfops.conv2d_same(img.double(), w.permute(1, 2, 3, 0).double(), 2)
  got = tfops.conv2d(packed.float().cpu().double(), w2.permute(1, 2, 3, 0).double(), 1, 1, (2, 1, 0, 0))
  assert got.shape == ref.shape
  assert float((got - ref).abs().max()) < 1e-9


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_zero_insert_and_batched_flip(cuda, dtype):
  """Bit-exact data movement: stride-2 zero insertion (strided dgrad on the tensor cores) and the
  one-launch refresh of several dgrad filter banks."""
  from wlseg import ops
  g = torch.Generator().manual_seed(3)
  src = torch.randn((2, 5, 7, 16), generator=g).to(dtype)
  dst = torch.full((2, 9, 13, 16), 7.0, dtype=dtype, device=cuda)
  ops.zero_insert(src.to(cuda), dst, 2)
  want = torch.zeros((2, 9, 13, 16), dtype=dtype)
  want[:, ::2, ::2, :] = src
  assert torch.equal(dst.cpu(), want)
  shapes = [(8, 3, 3, 16), (24, 1, 1, 8), (16, 4, 1, 8)]
  offs, n = [], 0
  for s in shapes:
    offs.append(n)
    n += s[0] * s[1] * s[2] * s[3]
  arena = torch.randn(n, generator=g).to(dtype)
  out = torch.zeros(n, dtype=dtype, device=cuda)
  table = torch.tensor([[o, o, *s] for o, s in zip(offs, shapes)], dtype=torch.int32, device=cuda)
  ops.weights_transpose_flip_batched(arena.to(cuda), out, table)
  for o, (K, R, S, C) in zip(offs, shapes):
    w = arena[o:o + K * R * S * C].view(K, R, S, C)
    ref = w.flip(1, 2).permute(3, 1, 2, 0).contiguous()
    assert torch.equal(out[o:o + K * R * S * C].view(C, R, S, K).cpu(), ref)


def test_weak_label_rasteriser_bit_exact_with_generate_rla(cuda):
  """Device-side box rasterisation + per-pixel normalisation (csrc/weak_labels.cu) against the numpy
  restatement of `_generate_rla` (input_subset_bboxes_v2.py:74-98) - bit-exact - including the worked
  examples of its comment (:87-95): [1,0,0]->[1,0,0], [2,0,0]->[1,0,0], [1,1,0]->[1/2,1/2,0],
  [2,1,0]->[2/3,1/3,0], no box -> void = 1; every pixel sums to 1 (input_subset_bboxes_v2_test.py:40-43)."""
  from oracle import weak_labels as oweak
  from wlseg import ops
  H, W = 37, 53
  g = torch.Generator().manual_seed(21)
  lists = []
  # image 0: the comment's cases laid out as overlapping boxes of classes 0 (car), 1 (bus)
  lists.append([(0, 0.0, 0.2, 0.0, 0.2),                                  # [1,0,0]
                (0, 0.3, 0.5, 0.0, 0.2), (0, 0.3, 0.5, 0.0, 0.2),         # [2,0,0]
                (0, 0.6, 0.8, 0.0, 0.2), (1, 0.6, 0.8, 0.0, 0.2),         # [1,1,0]
                (0, 0.0, 0.2, 0.5, 0.7), (0, 0.0, 0.2, 0.5, 0.7), (1, 0.0, 0.2, 0.5, 0.7)])  # [2,1,0]
  # image 1: random boxes incl. degenerate ones, full-image box, xmax = 1.0 (stop index past the border)
  k = 40
  c = torch.rand(k, 4, generator=g)
  boxes = [(int(torch.randint(0, 14, (1,), generator=g)), float(min(a, b)), float(max(a, b)), float(min(cc, d)), float(max(cc, d)))
           for a, b, cc, d in c.tolist()]
  boxes += [(3, 0.0, 1.0, 0.0, 1.0), (5, 0.5, 0.5, 0.2, 0.9), (7, 0.99, 1.0, 0.99, 1.0)]
  lists.append(boxes)
  lists.append([])   # image 2: no box at all -> void everywhere
  B = max(len(b) for b in lists)
  coords = torch.zeros(len(lists), B, 4)
  cids = torch.full((len(lists), B), -1, dtype=torch.int32)
  for i, bl in enumerate(lists):
    for j, (cid, x0, x1, y0, y1) in enumerate(bl):
      coords[i, j] = torch.tensor([x0, x1, y0, y1])
      cids[i, j] = cid
  got = ops.rasterize_bbox_labels(coords.to(cuda), cids.to(cuda), H, W).cpu()
  for i, bl in enumerate(lists):
    # the oracle sees the float32-rounded coordinates the device sees
    bl32 = [(cid, *[float(v) for v in coords[i, j]]) for j, (cid, *_r) in enumerate(bl)]
    want = torch.from_numpy(oweak.bbox_labels(bl32, H, W))
    assert torch.equal(got[i], want), f'image {i}'
  assert float((got.sum(-1) - 1.0).abs().max()) < 1e-3
  px = got[0, int(0.1 * H), :, :3]
  assert px[int(0.1 * W)].tolist() == [1.0, 0.0, 0.0] and px[int(0.4 * W)].tolist() == [1.0, 0.0, 0.0]
  assert px[int(0.7 * W)].tolist() == [0.5, 0.5, 0.0]
  q = got[0, int(0.6 * H), int(0.1 * W), :3].tolist()
  assert q == [float(np.float32(2) / np.float32(3)), float(np.float32(1) / np.float32(3)), 0.0]
  assert got[2, ..., 14].min() == 1.0 and got[2, ..., :14].abs().max() == 0.0
  # image-level labels: the vector tiled over the image
  vec = torch.zeros(2, 15)
  vec[0, [2, 9]] = 0.5
  vec[1, 14] = 1.0
  tiled = ops.tile_image_labels(vec.to(cuda), H, W).cpu()
  assert torch.equal(tiled[0], torch.from_numpy(oweak.image_labels([2, 9], H, W)))
  assert torch.equal(tiled[1], torch.from_numpy(oweak.image_labels([], H, W)))
