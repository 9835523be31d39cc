"""Parity of the convolution kernels with the oracle's conv (TF SAME / conv2d_same semantics).

fp32 direct kernel: 1e-4 relative to max|ref| (north star check mode).
bf16 kernels (direct and tcgen05 implicit GEMM): inputs are rounded to bf16 first, the oracle
convolves the rounded values in fp32, outputs must agree within 1e-2 relative to max|ref| (one bf16
rounding of the output + fp32 accumulation-order noise).
"""

import pytest
import torch

from oracle import tfops

pytestmark = pytest.mark.gpu


def _ref_conv(x, w_krsc, stride, dilation, scale, shift, res, res_stride, relu):
  w = w_krsc.permute(1, 2, 3, 0)  # KRSC -> HWIO
  y = tfops.conv2d_same(x, w, stride, dilation)
  if scale is not None:
    y = y * scale + shift
  if res is not None:
    y = y + res[:, ::res_stride, ::res_stride][:, :y.shape[1], :y.shape[2]]
  return torch.relu(y) if relu else y


def _run_conv(cuda, N, H, W, C, K, R, stride, dilation, dtype, algo, with_epi=True, res_stride=1, relu=True, seed=0,
              y_dtype=None):
  from wlseg import arch, ops
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(N, H, W, C, generator=g).to(dtype).float()
  w = (torch.randn(K, R, R, C, generator=g) / (R * R * C) ** 0.5).to(dtype).float()
  pt, P = arch.same_pad_before(R, stride, dilation, H)
  pl, Q = arch.same_pad_before(R, stride, dilation, W)
  scale = shift = res = None
  if with_epi:
    scale = 0.5 + torch.rand(K, generator=g)
    shift = torch.randn(K, generator=g) * 0.2
    res = torch.randn(N, P * res_stride, Q * res_stride, K, generator=g).to(dtype).float()
  ref = _ref_conv(x, w, stride, dilation, scale, shift, res, res_stride, relu)
  assert ref.shape == (N, P, Q, K)
  code = ops.dtype_code(dtype)
  ydt = dtype if y_dtype is None else y_dtype
  y = torch.full((N, P, Q, K), float('nan'), dtype=ydt, device=cuda)
  resd = None if res is None else res.to(dtype).to(cuda)
  prm = ops.conv_params((N, H, W, C), (K, R, R, C), stride=stride, dilation=dilation, pad=(pt, pl), out_hw=(P, Q),
                        relu=relu, dtype=code, y_dtype=ops.dtype_code(ydt), algo=algo, res=resd, res_stride=res_stride)
  ops.conv2d_fprop(prm, x.to(dtype).to(cuda), w.to(dtype).to(cuda), y,
                   None if scale is None else scale.to(cuda), None if shift is None else shift.to(cuda), resd)
  torch.cuda.synchronize()
  return y.float().cpu(), ref


DIRECT_CASES = [
    # N, H, W, C, K, R, stride, dilation
    (1, 9, 11, 3, 8, 7, 2, 1),     # root conv shape class (C=3, 7x7/2)
    (2, 8, 10, 16, 24, 3, 1, 1),
    (1, 12, 12, 8, 8, 3, 1, 2),    # dilated
    (1, 13, 9, 8, 16, 3, 2, 1),    # strided 3x3 on odd sizes
    (1, 6, 6, 32, 14, 1, 1, 1),    # logits-like 1x1
]


@pytest.mark.parametrize('case', DIRECT_CASES)
def test_conv_direct_fp32(cuda, case):
  from wlseg import ops
  got, ref = _run_conv(cuda, *case, dtype=torch.float32, algo=ops.ALGO_DIRECT)
  assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize('case', DIRECT_CASES[1:])
def test_conv_direct_bf16(cuda, case):
  from wlseg import ops
  got, ref = _run_conv(cuda, *case, dtype=torch.bfloat16, algo=ops.ALGO_DIRECT)
  assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())


TC_CASES = [
    # N, H, W, C, K, R, stride, dilation
    (1, 8, 16, 64, 64, 1, 1, 1),       # one tile, one K block
    (2, 16, 32, 64, 64, 3, 1, 1),      # 3x3: TMA zero-fill is the padding
    (1, 16, 16, 128, 256, 1, 1, 1),    # BN = 256
    (1, 24, 40, 256, 128, 3, 1, 2),    # dilation 2 (block3)
    (1, 16, 24, 128, 128, 3, 1, 4),    # dilation 4 (block4)
    (1, 13, 19, 64, 96, 3, 1, 1),      # ragged spatial size, K not a multiple of the tile
    (1, 12, 20, 32, 24, 1, 1, 1),      # C < 64 (zero-filled K block), K = 24 (BN = 32, scalar stores)
    (1, 32, 32, 64, 64, 3, 2, 1),      # stride 2 via TMA element strides (block1/unit_3/conv2)
    (1, 17, 23, 64, 64, 3, 2, 1),      # stride 2, odd sizes
    (3, 8, 16, 512, 2048, 1, 1, 1),    # many N tiles, deep K
    (1, 4, 4, 64, 64, 3, 1, 1),        # tiny image (TW = 4)
    (1, 1, 200, 64, 64, 1, 1, 1),      # P == 1 (TW = 128)
]


@pytest.mark.parametrize('case', TC_CASES)
def test_conv_tcgen05_bf16(cuda, case):
  from wlseg import ops
  got, ref = _run_conv(cuda, *case, dtype=torch.bfloat16, algo=ops.ALGO_TCGEN05, seed=sum(case))
  assert not torch.isnan(got).any(), 'some outputs were never written'
  assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())


PAIR_CASES = [
    # N, H, W, C, K, R, stride, dilation -- every one a 256-wide N tile with the staged epilogue
    (1, 16, 16, 128, 256, 1, 1, 1),    # one pair of M tiles, 2 k-blocks, residual
    (3, 8, 16, 512, 2048, 1, 1, 1),    # ODD M-tile count: the last pair's second CTA works on a phantom tile
    (2, 16, 24, 256, 512, 3, 1, 2),    # dilated 3x3, two N tiles
    (1, 13, 19, 256, 256, 3, 1, 1),    # ragged spatial size (clipped stores, zero-filled loads)
    (1, 40, 40, 1024, 768, 1, 1, 1),   # three N tiles, 13 M tiles
]


@pytest.mark.parametrize('pair', ['0', '1'])
@pytest.mark.parametrize('case', PAIR_CASES)
def test_conv_tcgen05_cta_pair_forced(cuda, case, pair, monkeypatch):
  """The CTA-pair (cta_group::2) form and the single-CTA form of the same layers, each forced through WLSEG_PAIR,
  against the oracle convolution (with folded BN, residual and ReLU): one bf16 output rounding."""
  from wlseg import ops
  monkeypatch.setenv('WLSEG_PAIR', pair)
  got, ref = _run_conv(cuda, *case, dtype=torch.bfloat16, algo=ops.ALGO_TCGEN05, seed=sum(case))
  assert not torch.isnan(got).any(), 'some outputs were never written'
  assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
  got, ref = _run_conv(cuda, *case, dtype=torch.bfloat16, algo=ops.ALGO_TCGEN05, seed=sum(case), with_epi=False, relu=False)
  assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())


@pytest.mark.parametrize('pair', ['0', '1'])
def test_conv_tcgen05_fused_bn_statistics_cta_pair(cuda, pair, monkeypatch):
  """Fused sum / sum-of-squares in the pair form: K = 512 (two N tiles: the cluster count is a multiple of 2), an odd
  number of M tiles (the phantom tile must not reach the statistics) and a ragged image."""
  from wlseg import arch, ops
  monkeypatch.setenv('WLSEG_PAIR', pair)
  g = torch.Generator().manual_seed(9)
  N, H, W, C, K = 3, 13, 17, 128, 512
  x = torch.randn(N, H, W, C, generator=g).to(torch.bfloat16)
  w = (torch.randn(K, 3, 3, C, generator=g) / 34).to(torch.bfloat16)
  pt, P = arch.same_pad_before(3, 1, 1, H)
  y = torch.full((N, H, W, K), float('nan'), dtype=torch.bfloat16, device=cuda)
  s1 = torch.zeros(K, dtype=torch.float64, device=cuda)
  s2 = torch.zeros(K, dtype=torch.float64, device=cuda)
  prm = ops.conv_params((N, H, W, C), (K, 3, 3, C), pad=(pt, pt), out_hw=(H, W), dtype=ops.BF16, algo=ops.ALGO_TCGEN05)
  ops.conv2d_fprop(prm, x.to(cuda), w.to(cuda), y, bn_sum=s1, bn_sqsum=s2)
  torch.cuda.synchronize()
  ref = tfops.conv2d_same(x.float(), w.float().permute(1, 2, 3, 0), 1, 1)
  assert not torch.isnan(y).any()
  assert float((y.float().cpu() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
  yq = y.float().cpu().double()   # the statistics are those of the STORED bf16 tensor
  assert torch.allclose(s1.cpu(), yq.sum((0, 1, 2)), rtol=1e-4, atol=1e-2)
  assert torch.allclose(s2.cpu(), (yq ** 2).sum((0, 1, 2)), rtol=1e-4, atol=1e-2)


def test_conv_tcgen05_plain_and_strided_residual(cuda):
  from wlseg import ops
  got, ref = _run_conv(cuda, 1, 16, 32, 64, 128, 1, 1, 1, dtype=torch.bfloat16, algo=ops.ALGO_TCGEN05,
                       with_epi=False, relu=False)
  assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
  got, ref = _run_conv(cuda, 1, 16, 32, 64, 128, 1, 1, 1, dtype=torch.bfloat16, algo=ops.ALGO_TCGEN05,
                       res_stride=2)
  assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())


def test_conv_tcgen05_fp32_output(cuda):
  from wlseg import ops
  got, ref = _run_conv(cuda, 1, 8, 16, 256, 14, 1, 1, 1, dtype=torch.bfloat16, algo=ops.ALGO_TCGEN05,
                       with_epi=False, relu=False, y_dtype=torch.float32)
  assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


def test_conv_tcgen05_fused_bn_statistics(cuda):
  from wlseg import arch, ops
  g = torch.Generator().manual_seed(5)
  N, H, W, C, K = 2, 13, 17, 64, 96
  x = torch.randn(N, H, W, C, generator=g).to(torch.bfloat16)
  w = (torch.randn(K, 3, 3, C, generator=g) / 24).to(torch.bfloat16)
  pt, P = arch.same_pad_before(3, 1, 1, H)
  y = torch.empty((N, H, W, K), dtype=torch.bfloat16, device=cuda)
  s1 = torch.zeros(K, dtype=torch.float64, device=cuda)
  s2 = torch.zeros(K, dtype=torch.float64, device=cuda)
  prm = ops.conv_params((N, H, W, C), (K, 3, 3, C), pad=(pt, pt), out_hw=(H, W), dtype=ops.BF16, algo=ops.ALGO_TCGEN05)
  ops.conv2d_fprop(prm, x.to(cuda), w.to(cuda), y, bn_sum=s1, bn_sqsum=s2)
  torch.cuda.synchronize()
  ref = tfops.conv2d_same(x.float(), w.float().permute(1, 2, 3, 0), 1, 1)
  assert torch.allclose(s1.float().cpu(), ref.sum((0, 1, 2)), rtol=1e-3, atol=1e-2)
  assert torch.allclose(s2.float().cpu(), (ref ** 2).sum((0, 1, 2)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize('case', [(2, 8, 10, 16, 24, 3, 1, 1), (1, 13, 9, 8, 16, 3, 2, 1), (1, 12, 12, 8, 8, 3, 1, 2),
                                  (1, 9, 11, 3, 8, 7, 2, 1)])
def test_conv_direct_backward_fp32(cuda, case):
  from wlseg import arch, ops
  N, H, W, C, K, R, stride, dilation = case
  g = torch.Generator().manual_seed(4)
  x = torch.randn(N, H, W, C, generator=g, requires_grad=True)
  w = (torch.randn(K, R, R, C, generator=g) / (R * R * C) ** 0.5).requires_grad_(True)
  y = tfops.conv2d_same(x, w.permute(1, 2, 3, 0), stride, dilation)
  dy = torch.randn(y.shape, generator=g)
  y.backward(dy)
  pt, P = arch.same_pad_before(R, stride, dilation, H)
  pl, Q = arch.same_pad_before(R, stride, dilation, W)
  prm = ops.conv_params((N, H, W, C), (K, R, R, C), stride=stride, dilation=dilation, pad=(pt, pl), out_hw=(P, Q),
                        dtype=ops.F32, algo=ops.ALGO_DIRECT)
  dx = torch.full((N, H, W, C), float('nan'), device=cuda)
  dw = torch.full((K, R, R, C), float('nan'), device=cuda)
  ops.conv2d_dgrad(prm, dy.to(cuda), w.detach().to(cuda), dx)
  ops.conv2d_wgrad(prm, x.detach().to(cuda), dy.to(cuda), dw)
  torch.cuda.synchronize()
  assert float((dx.cpu() - x.grad).abs().max()) <= 1e-4 * float(x.grad.abs().max())
  assert float((dw.cpu() - w.grad).abs().max()) <= 1e-4 * float(w.grad.abs().max())


def test_dgrad_as_flipped_fprop_tcgen05(cuda):
  """Stride-1 dgrad runs on the tensor cores as an fprop over dy with the rotated, transposed bank."""
  from wlseg import arch, ops
  g = torch.Generator().manual_seed(8)
  N, H, W, C, K, R, dil = 1, 16, 24, 64, 128, 3, 2
  x = torch.randn(N, H, W, C, generator=g, requires_grad=True)
  w = (torch.randn(K, R, R, C, generator=g) / 24).to(torch.bfloat16).float().requires_grad_(True)
  y = tfops.conv2d_same(x, w.permute(1, 2, 3, 0), 1, dil)
  dy = torch.randn(y.shape, generator=g).to(torch.bfloat16).float()
  y.backward(dy)
  wf = torch.empty((C, R, R, K), dtype=torch.bfloat16, device=cuda)
  ops.weights_transpose_flip(w.detach().to(torch.bfloat16).to(cuda), wf)
  pad = dil * (R - 1) - arch.same_pad_before(R, 1, dil, H)[0]
  prm = ops.conv_params((N, H, W, K), (C, R, R, K), dilation=dil, pad=(pad, pad), out_hw=(H, W), dtype=ops.BF16,
                        algo=ops.ALGO_TCGEN05)
  dx = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=cuda)
  ops.conv2d_fprop(prm, dy.to(torch.bfloat16).to(cuda), wf, dx)
  torch.cuda.synchronize()
  assert float((dx.float().cpu() - x.grad).abs().max()) <= 1e-2 * float(x.grad.abs().max())


WGRAD_TC_CASES = [
    # N, H, W, C, K, R, stride, dilation
    (1, 8, 8, 64, 64, 1, 1, 1),        # one patch, one chunk (BN = 64), K < 128 (zero-filled rows)
    (2, 16, 24, 128, 128, 1, 1, 1),    # BN = 128
    (1, 16, 16, 64, 128, 3, 1, 1),     # 9 taps x 1 chunk: N tiles of 4 chunks, the last one partial
    (2, 24, 40, 256, 256, 3, 1, 2),    # dilation 2 (block3), two M tiles, pixel splits
    (1, 16, 24, 128, 64, 3, 1, 4),     # dilation 4 (block4)
    (1, 13, 19, 64, 96, 3, 1, 1),      # ragged spatial size, K not a multiple of 64
    (1, 32, 32, 64, 64, 3, 2, 1),      # stride 2 (block1/unit_3/conv2)
    (1, 17, 23, 64, 64, 3, 2, 1),      # stride 2, odd sizes
    (3, 8, 16, 512, 256, 1, 1, 1),     # deep C: 8 chunks -> 2 N tiles
    (1, 12, 20, 256, 14, 1, 1, 1),     # logits layer: K = 14 (dy pitch padded to 16)
    (1, 1, 200, 64, 64, 1, 1, 1),      # P == 1
    (2, 16, 24, 256, 512, 1, 1, 1),    # CTA pair: two M pairs, one N tile
    (1, 24, 24, 512, 512, 3, 1, 4),    # CTA pair: dilated 3x3, 18 N tiles, pixel splits
    (1, 13, 19, 1024, 256, 1, 1, 1),   # CTA pair: deep C, ragged spatial size
]


@pytest.mark.parametrize('pair', ['0', '1'])
@pytest.mark.parametrize('case', WGRAD_TC_CASES)
def test_conv_wgrad_tcgen05(cuda, case, pair, monkeypatch):
  """dw on the tensor cores (MN-major operands, TMA reduce-add of the pixel splits) against autograd of
  the oracle convolution on the same bf16-rounded operands: 2e-3 of max|dw| (fp32 accumulation, bf16 inputs
  are exact in both)."""
  from wlseg import arch, ops
  N, H, W, C, K, R, stride, dilation = case
  eligible = K % 256 == 0 and C % 64 == 0 and (R * R * (C // 64)) % 4 == 0
  if pair == '0' and not eligible:
    pytest.skip('the single-CTA form is what the default run of this case covers')
  monkeypatch.setenv('WLSEG_WGRAD_PAIR', pair)   # '1' = default policy (CTA pair where eligible), '0' = never
  g = torch.Generator().manual_seed(sum(case))
  x = torch.randn(N, H, W, C, generator=g).to(torch.bfloat16).float().requires_grad_(True)
  w = (torch.randn(K, R, R, C, generator=g) / (R * R * C) ** 0.5).requires_grad_(True)
  y = tfops.conv2d_same(x, w.permute(1, 2, 3, 0), stride, dilation)
  dy = torch.randn(y.shape, generator=g).to(torch.bfloat16).float()
  y.backward(dy)
  pt, P = arch.same_pad_before(R, stride, dilation, H)
  pl, Q = arch.same_pad_before(R, stride, dilation, W)
  kp = (K + 7) // 8 * 8
  dyd = torch.zeros((N, P, Q, kp), dtype=torch.bfloat16, device=cuda)
  dyd[..., :K] = dy.to(torch.bfloat16).to(cuda)
  prm = ops.conv_params((N, H, W, C), (K, R, R, C), stride=stride, dilation=dilation, pad=(pt, pl), out_hw=(P, Q),
                        y_pitch=kp, dtype=ops.BF16, algo=ops.ALGO_TCGEN05)
  dw = torch.full((K, R, R, C), float('nan'), device=cuda)
  ops.conv2d_wgrad(prm, x.detach().to(torch.bfloat16).to(cuda), dyd, dw)
  torch.cuda.synchronize()
  assert not torch.isnan(dw).any()
  err = float((dw.cpu() - w.grad).abs().max()) / float(w.grad.abs().max())
  print(f'wgrad {case}: max-rel {err:.3e}')
  assert err <= 2e-3


BNB_CASES = [
    # N, H, W, C, K, R, dilation       (K = channels of the tensor whose gradient leaves: BN tile 64 / 128 / 256 / pair)
    (2, 24, 40, 64, 64, 3, 1),         # block1 conv2 dgrad, BN = 64
    (2, 20, 24, 512, 128, 1, 1),       # block2 conv3 dgrad, BN = 128
    (1, 19, 27, 128, 128, 3, 1),       # ragged spatial size: partial tiles must not reach the sums
    (2, 24, 24, 1024, 256, 1, 1),      # block3 conv3 dgrad, BN = 256 (CTA pair), 16 k-blocks
    (1, 16, 24, 256, 256, 3, 2),       # block3 conv2 dgrad, dilated
    (1, 17, 24, 512, 512, 3, 4),       # block4 conv2 dgrad, two N tiles
    (1, 24, 40, 256, 256, 1, 1),       # 15 M tiles: the last pair's second CTA runs a phantom tile
    (4, 96, 96, 256, 256, 3, 2),       # BASELINE size: 288 M tiles, 2 units per CTA pair (z ring + staging reuse across tiles)
    (2, 128, 136, 64, 64, 3, 1),       # 306 tiles on 148 single CTAs, ragged width
]


@pytest.mark.parametrize('case', BNB_CASES)
def test_fprop_bnbwd_equals_conv_then_bn_bwd_reduce(cuda, case):
  """wlseg_conv2d_fprop_bnbwd (the dgrad epilogue applies the ReLU derivative of the BN layer below and accumulates
  its dgamma / dbeta) against the two launches it replaces: wlseg_conv2d_fprop, then wlseg_bn_bwd_reduce with the
  mask recomputed from z.  The masked gradient must be BIT-IDENTICAL to mask(bf16(conv)) - rounding and masking
  commute - and the sums agree to fp32 summation-order noise (1e-5 of the sum of magnitudes)."""
  from wlseg import ops
  N, H, W, C, K, R, dil = case
  g = torch.Generator().manual_seed(sum(case))
  dt = torch.bfloat16
  x = torch.randn(N, H, W, C, generator=g).to(dt).to(cuda)
  w = (torch.randn(K, R, R, C, generator=g) / (R * R * C) ** 0.5).to(dt).to(cuda)
  z = torch.randn(N, H, W, K, generator=g).to(dt).to(cuda)
  scale = (0.5 + torch.rand(K, generator=g)).to(cuda)
  shift = (torch.randn(K, generator=g) * 0.3).to(cuda)
  mean = (torch.randn(K, generator=g) * 0.2).to(cuda)
  invstd = (0.5 + torch.rand(K, generator=g)).to(cuda)
  pad = dil * (R - 1) // 2
  count = N * H * W
  code = ops.dtype_code(dt)
  # unfused: convolution, then the reduction pass (mask from the sign of z * scale + shift)
  y0 = torch.empty((N, H, W, K), dtype=dt, device=cuda)
  prm0 = ops.conv_params((N, H, W, C), (K, R, R, C), dilation=dil, pad=(pad, pad), out_hw=(H, W), dtype=code)
  ops.conv2d_fprop(prm0, x, w, y0)
  dg0 = torch.zeros(K, dtype=torch.float64, device=cuda)
  db0 = torch.zeros(K, dtype=torch.float64, device=cuda)
  ops.bn_bwd_reduce(y0, None, z, mean, invstd, count, K, True, dg0, db0, scale=scale, shift=shift, pitch=K)
  # fmaf(z, scale, shift) is the correctly rounded value of the exact real number z * scale + shift, so its sign is the
  # sign of that number: evaluated here in fp64, where product and sum of these operands are exact
  on = (z.double() * scale.double() + shift.double()) > 0
  # fused
  y1 = torch.full((N, H, W, K), float('nan'), dtype=dt, device=cuda)
  dg1 = torch.zeros(K, dtype=torch.float64, device=cuda)
  db1 = torch.zeros(K, dtype=torch.float64, device=cuda)
  prm1 = ops.conv_params((N, H, W, C), (K, R, R, C), dilation=dil, pad=(pad, pad), out_hw=(H, W), dtype=code, res=z,
                         res_stride=1)
  ops.conv2d_fprop_bnbwd(prm1, x, w, y1, z, scale, shift, mean, invstd, dg1, db1)
  torch.cuda.synchronize()
  want = torch.where(on, y0, torch.zeros_like(y0))
  diff = (y1.float() - want.float()).abs()
  assert int((diff > 0).sum()) == 0, f'masked gradient differs at {int((diff > 0).sum())} elements (max {float(diff.max()):.3e})'
  mag_b = (want.float().abs().double()).sum((0, 1, 2))
  mag_g = (want.float().abs().double() * ((z.float() - mean).abs() * invstd).double()).sum((0, 1, 2))
  eb = float(((db1 - db0).abs() / (mag_b + 1e-12)).max())
  eg = float(((dg1 - dg0).abs() / (mag_g + 1e-12)).max())
  print(f'bnbwd {case}: dbeta err {eb:.2e}, dgamma err {eg:.2e} (relative to the sum of magnitudes)')
  assert eb <= 1e-5 and eg <= 1e-5


@pytest.mark.parametrize('case', [(2, 24, 40, 64, 64, 3, 1), (2, 20, 24, 128, 512, 1, 1), (4, 96, 96, 256, 256, 3, 2),
                                  (1, 17, 24, 512, 2048, 1, 1)])
def test_fprop_bn_equals_conv_then_bn_finalize(cuda, case):
  """wlseg_conv2d_fprop_bn (the last CTA of the convolution grid finalises the batch norm) against the two launches
  it replaces, wlseg_conv2d_fprop with fused statistics + wlseg_bn_finalize: same raw output bit for bit, scale / shift
  / saved mean / inverse std / moving statistics to 1e-6 (the fp64 atomics commit in another order); the ticket word
  is left at zero, so that consecutive layers share it."""
  from wlseg import ops
  N, H, W, C, K, R, dil = case
  g = torch.Generator().manual_seed(sum(case) + 1)
  dt = torch.bfloat16
  x = torch.randn(N, H, W, C, generator=g).to(dt).to(cuda)
  w = (torch.randn(K, R, R, C, generator=g) / (R * R * C) ** 0.5).to(dt).to(cuda)
  gamma = (0.5 + torch.rand(K, generator=g)).to(cuda)
  beta = (torch.randn(K, generator=g) * 0.3).to(cuda)
  pad = dil * (R - 1) // 2
  count = N * H * W
  prm = ops.conv_params((N, H, W, C), (K, R, R, C), dilation=dil, pad=(pad, pad), out_hw=(H, W), dtype=ops.dtype_code(dt))
  counter = torch.zeros(1, dtype=torch.int32, device=cuda)
  outs = []
  for fused in (False, True, True):     # the second fused call reuses the counter
    y = torch.full((N, H, W, K), float('nan'), dtype=dt, device=cuda)
    s1 = torch.zeros(K, dtype=torch.float64, device=cuda)
    s2 = torch.zeros(K, dtype=torch.float64, device=cuda)
    mm = torch.full((K,), 0.25, device=cuda)
    mv = torch.full((K,), 1.5, device=cuda)
    o = [torch.full((K,), float('nan'), device=cuda) for _ in range(4)]
    if fused:
      ops.conv2d_fprop_bn(prm, x, w, y, s1, s2, count, gamma, beta, 1e-5, 0.9, mm, mv, o[0], o[1], o[2], o[3], counter)
    else:
      ops.conv2d_fprop(prm, x, w, y, bn_sum=s1, bn_sqsum=s2)
      ops.bn_finalize(s1, s2, count, K, gamma, beta, 1e-5, 0.9, mm, mv, o[0], o[1], o[2], o[3])
    torch.cuda.synchronize()
    assert int(counter.item()) == 0
    outs.append((y.cpu(), [t.cpu() for t in o + [mm, mv]]))
  for y, vec in outs[1:]:
    assert torch.equal(y, outs[0][0])
    for name, a, b in zip(('scale', 'shift', 'mean', 'invstd', 'moving_mean', 'moving_var'), vec, outs[0][1]):
      err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
      assert err <= 1e-6, f'{name}: {err:.2e}'


HALO_CASES = [
    # N, H, W, R, S, pad_top, pad_left, epilogue      (C = K = 64, stride 1: the halo form of conv_igemm_kernel)
    (2, 40, 56, 3, 3, 1, 1, 'plain'),
    (1, 37, 45, 3, 3, 1, 1, 'bn_relu_res'),     # ragged tiles on both axes, folded BN + residual + ReLU
    (2, 64, 96, 3, 3, 1, 1, 'stats'),           # training forward: raw output + fused statistics
    (1, 50, 64, 4, 1, 2, 0, 'bn_relu'),         # the packed root convolution: R = 4, S = 1, pad (2, 0)
    (4, 192, 192, 3, 3, 1, 1, 'plain'),         # BASELINE block1 size: 1152 tiles, 8 per CTA (patch double buffering)
    (2, 128, 256, 4, 1, 2, 0, 'stats'),
]


@pytest.mark.parametrize('case', HALO_CASES)
def test_halo_form_equals_per_tap_form_and_oracle(cuda, case, monkeypatch):
  """The halo form (one activation patch per tile, taps as shifted shared-memory descriptors, resident filters) against
  the per-tap form of the same kernel (WLSEG_HALO=0) - same MMAs in the same order: BIT-IDENTICAL outputs - and against
  the oracle's convolution (1e-2 of max|ref|: one bf16 output rounding)."""
  from wlseg import ops
  N, H, W, R, S, pt, pl, epi = case
  C = K = 64
  g = torch.Generator().manual_seed(N * 1000 + H + W + R)
  dt = torch.bfloat16
  x = torch.randn(N, H, W, C, generator=g).to(dt)
  w = (torch.randn(K, R, S, C, generator=g) / (R * S * C) ** 0.5).to(dt)
  scale = shift = res = None
  relu = False
  if epi.startswith('bn_relu'):
    scale = 0.5 + torch.rand(K, generator=g)
    shift = torch.randn(K, generator=g) * 0.2
    relu = True
  if epi == 'bn_relu_res':
    res = torch.randn(N, H, W, K, generator=g).to(dt)
  P, Q = H, W     # SAME geometry: R = 3 pads (1, 1); the packed root R = 4 pads 2 above and 1 below
  prm = ops.conv_params((N, H, W, C), (K, R, S, C), pad=(pt, pl), out_hw=(P, Q), relu=relu, dtype=ops.dtype_code(dt),
                        res=None if res is None else res.to(cuda), res_stride=1)
  outs = []
  for halo, split in (('1', '1'), ('0', '0'), ('0', '1'), ('1', '0')):
    monkeypatch.setenv('WLSEG_HALO', halo)
    monkeypatch.setenv('WLSEG_EPI_SPLIT', split)    # BN = 64: the epilogue column groups alternate tiles
    y = torch.full((N, P, Q, K), float('nan'), dtype=dt, device=cuda)
    s1 = torch.zeros(K, dtype=torch.float64, device=cuda) if epi == 'stats' else None
    s2 = torch.zeros(K, dtype=torch.float64, device=cuda) if epi == 'stats' else None
    ops.conv2d_fprop(prm, x.to(cuda), w.to(cuda), y, None if scale is None else scale.to(cuda),
                     None if shift is None else shift.to(cuda), None if res is None else res.to(cuda), s1, s2)
    torch.cuda.synchronize()
    outs.append((y.cpu(), None if s1 is None else (s1.cpu(), s2.cpu())))
  (y1, st1), (y0, st0) = outs[0], outs[1]
  for tag, (yv, stv) in zip(('halo + split', 'per-tap + split', 'halo'), (outs[0], outs[2], outs[3])):
    assert torch.equal(yv, y0), (f'{tag} vs per-tap: {int((yv != y0).sum())} elements differ, max '
                                 f'{float((yv.float() - y0.float()).abs().max()):.3e}')
    if stv is not None:
      for a, b in zip(stv, st0):
        assert float((a - b).abs().max() / b.abs().max()) <= 1e-6
  # oracle: zero-pad explicitly, VALID correlation
  xp = torch.nn.functional.pad(x.float().permute(0, 3, 1, 2), (pl, Q + S - 1 - W - pl, pt, P + R - 1 - H - pt))
  ref = torch.nn.functional.conv2d(xp, w.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
  if scale is not None:
    ref = ref * scale + shift
  if res is not None:
    ref = ref + res.float()
  if relu:
    ref = torch.relu(ref)
  err = float((y1.float() - ref).abs().max() / ref.abs().max())
  assert err <= 1e-2, err
