"""Parameters and the forward / backward drivers of the network on libwlseg kernels.

Host-side mirror of the reference's `model()` (code/models/resnet50_extended_model_hierarchical.py:17-141)
and `feature_extractor()` (code/models/resnet50_extended_feature_extractor.py:8-51).  This file
only sequences kernel launches and owns buffers; all arithmetic is in csrc/.

Data layout in HBM
  activations  NHWC, bf16 (fp32 in the check mode)
  conv kernels KRSC; ONE fp32 arena holds [all conv kernels | all gammas | all betas] (the
               optimizer updates it with one launch), a parallel bf16 arena holds the MMA operands
  BN moving statistics: a separate fp32 arena [all means | all variances]
  low-res logits: fp32 [N, h, w, logits_pitch] with the three heads concatenated along channels
"""

import math
import os

import torch

from wlseg import arch, ops


def _trunc_normal_(t, std, gen):
  torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)


class Params:
  """All trainable variables + BN moving statistics, addressable by TF variable name."""

  def __init__(self, hier, device, output_stride=8, psp=False, fov=None, upsampling='bilinear', norm='batch'):
    self.hier = hier
    self.device = torch.device(device)
    self.psp = bool(psp)  # --psp_module: five more convolutions (arch.PSP_SCOPES)
    # --fov_expansion_kernel_size / --fov_expansion_kernel_rate: one dilated convolution (arch.FOV_SCOPE)
    self.fov = tuple(fov) if fov and fov[0] > 0 and fov[1] > 0 else None
    # --upsampling_method: 'bilinear' | 'no' (predictions stay at the feature resolution) | 'hybrid' (three more
    # layers, arch.UPSAMPLING_SCOPES: transposed convolutions with a bias and no batch norm, kept in fp32)
    if upsampling not in ('bilinear', 'no', 'hybrid'):
      raise ValueError('No such upsampling method.')   # models/resnet50_extended_model_hierarchical.py:181-182
    self.upsampling = upsampling
    self.specs = arch.conv_specs(hier.head_widths, output_stride, psp=self.psp, fov=self.fov, upsampling=upsampling)
    self.plain = set(arch.UPSAMPLING_SCOPES) if upsampling == 'hybrid' else set()   # layers without batch norm
    # --norm_layer: 'batch' | 'group' (tf.contrib.layers.group_norm: variables <scope>/GroupNorm/{beta,gamma}, no moving
    # statistics; 32 groups, 1 for the logits layers - models/resnet50_extended_model_hierarchical.py:75-77,314-333)
    if norm not in ('batch', 'group'):
      raise ValueError('norm_type not valid.')   # module_arg_scope, :296-297
    self.norm = norm
    self.norm_scope = 'BatchNorm' if norm == 'batch' else 'GroupNorm'
    self.by_scope = {s.scope: s for s in self.specs}
    self.w_off, self.c_off = {}, {}
    off = 0
    for s in self.specs:
      self.w_off[s.scope] = off
      off += s.K * s.R * s.S * s.C
    self.n_conv = off
    coff = 0
    for s in self.specs:
      self.c_off[s.scope] = coff
      coff += s.K
    self.n_chan = coff
    # arenas are padded to a multiple of 8 elements so that every region stays 16-byte aligned
    self.n_conv_pad = (self.n_conv + 7) // 8 * 8
    self.n_chan_pad = (self.n_chan + 7) // 8 * 8
    self.n_total = self.n_conv_pad + 2 * self.n_chan_pad
    dev = self.device
    self.master = torch.zeros(self.n_total, dtype=torch.float32, device=dev)
    self.operand = torch.zeros(self.n_total, dtype=torch.bfloat16, device=dev)
    self.moving = torch.zeros(2 * self.n_chan_pad, dtype=torch.float32, device=dev)
    self.moving[self.n_chan_pad:] = 1.0
    self.master[self.n_conv_pad:self.n_conv_pad + self.n_chan_pad] = 1.0  # gamma
    self._derived = {}
    self._pack_idx = None
    self.version = 0

  # ---- views -------------------------------------------------------------------------------
  def _wview(self, arena, scope):
    s = self.by_scope[scope]
    o = self.w_off[scope]
    return arena[o:o + s.K * s.R * s.S * s.C].view(s.K, s.R, s.S, s.C)

  def w32(self, scope):
    return self._wview(self.master, scope)

  def wbf(self, scope):
    return self._wview(self.operand, scope)

  def _cview(self, arena, base, scope, n=None):
    o = base + self.c_off[scope]
    return arena[o:o + (self.by_scope[scope].K if n is None else n)]

  def gamma(self, scope, n=None):
    return self._cview(self.master, self.n_conv_pad, scope, n)

  def beta(self, scope, n=None):
    return self._cview(self.master, self.n_conv_pad + self.n_chan_pad, scope, n)

  def moving_mean(self, scope, n=None):
    return self._cview(self.moving, 0, scope, n)

  def moving_var(self, scope, n=None):
    return self._cview(self.moving, self.n_chan_pad, scope, n)

  def groups(self, scope, K=None):
    """Number of normalisation groups of the layer `scope` covering K channels (K > the layer's own width for the
    merged adaptation conv1 layers: every branch keeps its own 32 groups)."""
    width = self.by_scope[scope].K
    gs = width if scope.startswith('softmax_classifier/') else width // 32
    return (width if K is None else K) // gs

  # ---- initialisation / import --------------------------------------------------------------
  def init_random(self, seed=0):
    """variance_scaling_initializer() defaults: truncated normal, stddev sqrt(1.3*2/fan_in)
    (code/models/resnet50_extended_model_hierarchical.py:337); gamma 1, beta 0, moving 0 / 1."""
    gen = torch.Generator().manual_seed(seed)
    host = torch.zeros(self.n_total, dtype=torch.float32)
    for s in self.specs:
      w = torch.empty(s.K, s.R, s.S, s.C)
      _trunc_normal_(w, math.sqrt(1.3 * 2.0 / (s.R * s.S * s.C)), gen)
      o = self.w_off[s.scope]
      host[o:o + w.numel()] = w.reshape(-1)
    host[self.n_conv_pad:self.n_conv_pad + self.n_chan_pad] = 1.0
    self.master.copy_(host)
    self.moving[:self.n_chan_pad] = 0.0
    self.moving[self.n_chan_pad:] = 1.0
    self.sync_operands()

  def load_tf_dict(self, tensors):
    """Import a {TF variable name: tensor} dict (conv kernels HWIO as TF stores them)."""
    host = self.master.cpu()
    mov = self.moving.cpu()
    for s in self.specs:
      w = tensors[f'{s.scope}/weights'].to(torch.float32)
      o = self.w_off[s.scope]
      c = self.c_off[s.scope]
      if s.scope in self.plain:
        # conv2d_transpose filter [kh, kw, out, in] -> the stride-1 correlation kernel it equals:
        # W[o, r, s, i] = f[kh-1-r, kw-1-s, o, i]; its bias lives in the beta slot (gamma stays 1, unused)
        assert tuple(w.shape) == (s.R, s.S, s.K, s.C), (s.scope, tuple(w.shape))
        host[o:o + w.numel()] = w.flip(0, 1).permute(2, 0, 1, 3).reshape(-1)
        host[self.n_conv_pad + self.n_chan_pad + c:self.n_conv_pad + self.n_chan_pad + c + s.K] = tensors[f'{s.scope}/biases']
        continue
      assert tuple(w.shape) == (s.R, s.S, s.C, s.K), (s.scope, tuple(w.shape))
      host[o:o + w.numel()] = w.permute(3, 0, 1, 2).reshape(-1)
      ns = self.norm_scope
      host[self.n_conv_pad + c:self.n_conv_pad + c + s.K] = tensors[f'{s.scope}/{ns}/gamma']
      host[self.n_conv_pad + self.n_chan_pad + c:self.n_conv_pad + self.n_chan_pad + c + s.K] = \
          tensors[f'{s.scope}/{ns}/beta']
      if self.norm == 'batch':
        mov[c:c + s.K] = tensors[f'{s.scope}/BatchNorm/moving_mean']
        mov[self.n_chan_pad + c:self.n_chan_pad + c + s.K] = tensors[f'{s.scope}/BatchNorm/moving_variance']
    self.master.copy_(host)
    self.moving.copy_(mov)
    self.sync_operands()

  def to_tf_dict(self):
    out = {}
    for s in self.specs:
      if s.scope in self.plain:
        out[f'{s.scope}/weights'] = self.w32(s.scope).permute(1, 2, 0, 3).flip(0, 1).contiguous().cpu()
        out[f'{s.scope}/biases'] = self.beta(s.scope).cpu().clone()
        continue
      out[f'{s.scope}/weights'] = self.w32(s.scope).permute(1, 2, 3, 0).contiguous().cpu()
      out[f'{s.scope}/{self.norm_scope}/gamma'] = self.gamma(s.scope).cpu().clone()
      out[f'{s.scope}/{self.norm_scope}/beta'] = self.beta(s.scope).cpu().clone()
      if self.norm == 'batch':
        out[f'{s.scope}/BatchNorm/moving_mean'] = self.moving_mean(s.scope).cpu().clone()
        out[f'{s.scope}/BatchNorm/moving_variance'] = self.moving_var(s.scope).cpu().clone()
    return out

  def arena_to_tf_dict(self, arena):
    """An arena with the master layout (momentum, EMA shadows, gradients) as {model variable name: tensor},
    conv kernels back in TF's HWIO layout (wlseg/checkpoints.py names the slots)."""
    out = {}
    for s in self.specs:
      if s.scope in self.plain:
        out[f'{s.scope}/weights'] = self._wview(arena, s.scope).permute(1, 2, 0, 3).flip(0, 1).contiguous().cpu()
        out[f'{s.scope}/biases'] = self._cview(arena, self.n_conv_pad + self.n_chan_pad, s.scope).cpu().clone()
        continue
      out[f'{s.scope}/weights'] = self._wview(arena, s.scope).permute(1, 2, 3, 0).contiguous().cpu()
      out[f'{s.scope}/{self.norm_scope}/gamma'] = self._cview(arena, self.n_conv_pad, s.scope).cpu().clone()
      out[f'{s.scope}/{self.norm_scope}/beta'] = self._cview(arena, self.n_conv_pad + self.n_chan_pad, s.scope).cpu().clone()
    return out

  def load_into_arena(self, arena, named):
    """Inverse of arena_to_tf_dict for the names present in `named`; the rest of the arena is kept."""
    host = arena.cpu()
    for s in self.specs:
      if s.scope in self.plain:
        t, b = named.get(f'{s.scope}/weights'), named.get(f'{s.scope}/biases')
        if t is not None:
          o = self.w_off[s.scope]
          host[o:o + t.numel()] = t.to(torch.float32).flip(0, 1).permute(2, 0, 1, 3).reshape(-1)
        if b is not None:
          base = self.n_conv_pad + self.n_chan_pad + self.c_off[s.scope]
          host[base:base + s.K] = b.to(torch.float32)
        continue
      t = named.get(f'{s.scope}/weights')
      if t is not None:
        assert tuple(t.shape) == (s.R, s.S, s.C, s.K), (s.scope, tuple(t.shape))
        o = self.w_off[s.scope]
        host[o:o + t.numel()] = t.to(torch.float32).permute(3, 0, 1, 2).reshape(-1)
      c = self.c_off[s.scope]
      for v, base in (('gamma', self.n_conv_pad), ('beta', self.n_conv_pad + self.n_chan_pad)):
        t = named.get(f'{s.scope}/{self.norm_scope}/{v}')
        if t is not None:
          host[base + c:base + c + s.K] = t.to(torch.float32)
    arena.copy_(host)

  def sync_operands(self):
    """bf16 operand copy of the master arena (the optimizer kernel keeps it in sync afterwards)."""
    ops.cast_f32_to_bf16(self.master, self.operand)
    self.touch()

  def touch(self):
    self.version += 1
    self._derived.clear()

  # ---- derived operands ----------------------------------------------------------------------
  def conv1_pack_index(self):
    """Column of the packed [64, 4*64] root filter bank that holds w[k, r, s, c] (flattened r, s, c):
    W2[k, a, 0, b*16 + (i*2+j)*3 + c] = w[k, 2a+i-1, 2b+j-1, c] (csrc/transform.cu)."""
    if self._pack_idx is None:
      ii = []
      for r in range(7):
        a, i = (r + 1) // 2, (r + 1) % 2
        for s_ in range(7):
          b, j = (s_ + 1) // 2, (s_ + 1) % 2
          for c in range(3):
            ii.append(a * 64 + b * 16 + (i * 2 + j) * 3 + c)
      self._pack_idx = torch.tensor(ii, dtype=torch.long, device=self.device)
    return self._pack_idx

  def conv1_packed_weights(self, dtype):
    """Root 7x7/2 kernel rewritten for the packed input of csrc/transform.cu (zero where the packed
    tap has no counterpart)."""
    key = ('conv1_packed', dtype)
    if key not in self._derived:
      w = self.w32(f'{arch.RES}/conv1').reshape(64, 147)
      w2 = torch.zeros(64, 256, dtype=torch.float32, device=self.device)
      w2[:, self.conv1_pack_index()] = w
      self._derived[key] = w2.to(dtype).view(64, 4, 1, 64)
    return self._derived[key]

  def folded_bn(self, scope, n=None, eps=1e-5):
    """Inference batch norm as per-channel scale / shift (fp32)."""
    key = ('fold', scope, n)
    if key not in self._derived:
      scale = self.gamma(scope, n) * torch.rsqrt(self.moving_var(scope, n) + eps)
      shift = self.beta(scope, n) - self.moving_mean(scope, n) * scale
      self._derived[key] = (scale.contiguous(), shift.contiguous())
    return self._derived[key]

  def flip_plan(self):
    """(table, arena, views) for the one-launch refresh of every stride-1/2 dgrad filter bank
    ([C, R, S, K], rotated 180 degrees) from the bf16 operand arena.  The three adaptation conv1
    kernels form ONE bank (they run as one 256 -> 768 GEMM); the root convolution has no input
    gradient; the logits layers (K = 14 / 7 / 3, padded to 8 / 16) keep their per-layer path."""
    if getattr(self, '_flip', None) is None:
      rows, views = [], {}
      arena = torch.zeros(self.n_total, dtype=torch.bfloat16, device=self.device)
      fused = [s.scope for s in self.specs if s.scope.startswith('adaptation_module/') and s.scope.endswith('/conv1')]
      for s in self.specs:
        if s.scope.endswith('resnet_v1_50/conv1') or s.K % 8 != 0:
          continue
        o = self.w_off[s.scope]
        if s.scope in fused:
          if s.scope != fused[0]:
            continue
          K = s.K * len(fused)
        else:
          K = s.K
        rows.append([o, o, K, s.R, s.S, s.C])
        views[s.scope] = arena[o:o + K * s.R * s.S * s.C].view(s.C, s.R, s.S, K)
      table = torch.tensor(rows, dtype=torch.int32, device=self.device)
      self._flip = (table, arena, views)
    return self._flip

  def flipped(self, scope, dtype):
    """Filter bank of the stride-1 dgrad-as-fprop: [C, R, S, K], rotated 180 degrees."""
    key = ('flip', scope, dtype)
    if key not in self._derived:
      src = self.wbf(scope) if dtype == torch.bfloat16 else self.w32(scope)
      s = self.by_scope[scope]
      dst = torch.empty(s.C, s.R, s.S, s.K, dtype=dtype, device=self.device)
      ops.weights_transpose_flip(src, dst)
      self._derived[key] = dst
    return self._derived[key]


class Network:
  """Sequences the kernels of one forward (and backward) pass.

  dtype=torch.bfloat16 is the product path (tcgen05 convolutions); dtype=torch.float32 is the
  check mode: same kernels for everything but the convolutions, which run the direct fp32 kernel.
  """

  def __init__(self, params, dtype=torch.bfloat16, bn_decay=0.9, eps=1e-5, conv_algo=ops.ALGO_AUTO):
    self.p = params
    self.hier = params.hier
    self.hstruct = params.hier.as_struct()
    self.dtype = dtype
    self.code = ops.dtype_code(dtype)
    self.bn_decay = bn_decay
    self.eps = eps
    self.conv_algo = conv_algo
    self.dev = params.device
    self.tape = None
    self.profile = None  # bench.py: list receiving one CUDA-event record per convolution launch
    # inference: consecutive convolutions may walk their tiles in alternating directions, so that a layer starts on
    # the rows its producer wrote last (wlseg_conv_params::reverse).  OFF by default: measured on B200 it LOSES for
    # the convolution chain (eval step 8.76 / 8.84 -> 8.85 / 9.01 ms), unlike the batch-norm passes of training,
    # which gain from the same idea (csrc/bn.cu); WLSEG_SNAKE=1 switches it on.
    import os
    self.snake = os.environ.get('WLSEG_SNAKE', '0') == '1'
    self._snake_flip = False

  # ---- helpers ---------------------------------------------------------------------------------
  class _Timed:
    """bench.py: CUDA-event pair around one launch of a bandwidth kernel (algorithmic bytes given)."""

    def __init__(self, net, cls, nbytes):
      self.net, self.rec = net, None
      if getattr(net, 'profile', None) is not None:
        self.rec = {'cls': cls, 'flops': 0.0, 'bytes': float(nbytes), 'sig': (cls,),
                    'e0': torch.cuda.Event(enable_timing=True), 'e1': torch.cuda.Event(enable_timing=True)}

    def __enter__(self):
      if self.rec is not None:
        self.rec['e0'].record()

    def __exit__(self, *exc):
      if self.rec is not None:
        self.rec['e1'].record()
        self.net.profile.append(self.rec)

  def _weights(self, scope):
    return self.p.wbf(scope) if self.dtype == torch.bfloat16 else self.p.w32(scope)

  def _geom(self, spec, H, W):
    pt, P = arch.same_pad_before(spec.R, spec.stride, spec.dilation, H)
    pl, Q = arch.same_pad_before(spec.S, spec.stride, spec.dilation, W)
    return pt, pl, P, Q

  def _conv(self, x, w, *, stride=1, dilation=1, pad=(0, 0), out_hw, scale=None, shift=None, relu=False,
            residual=None, res_stride=1, y=None, y_dtype=None, bn_sum=None, bn_sqsum=None, out_mask=None, bnb=None,
            fin=None):
    """bnb = (z, scale, shift, mean, invstd, dgamma, dbeta): the BN-backward form of a data gradient
    (ops.conv2d_fprop_bnbwd); z takes the residual slot of the parameters."""
    N, H, W, C = x.shape
    K = w.shape[0]
    if bnb is not None:
      assert residual is None and out_mask is None and scale is None and bn_sum is None and not relu
      residual = bnb[0]
    reverse = False
    if self.snake and self.tape is None and scale is not None and bn_sum is None:
      self._snake_flip = not self._snake_flip   # inference layers only (folded BN): alternate
      reverse = self._snake_flip
    if y is None:
      ydt = self.dtype if y_dtype is None else y_dtype
      y = torch.empty((N, out_hw[0], out_hw[1], K), dtype=ydt, device=self.dev)
    prm = ops.conv_params((N, H, W, C), tuple(w.shape), stride=stride, dilation=dilation, pad=pad, out_hw=out_hw,
                          x_pitch=x.stride(2), y_pitch=y.stride(2), relu=relu, dtype=self.code,
                          y_dtype=ops.dtype_code(y.dtype), algo=self.conv_algo, res=residual, res_stride=res_stride,
                          reverse=reverse)
    tc = self.conv_algo != ops.ALGO_DIRECT and ops.conv2d_tcgen05_supported(prm)
    rec = None
    if self.profile is not None:
      esz = 2 if self.dtype == torch.bfloat16 else 4
      rec = {'cls': ('igemm_bn%d' % (256 if K > 128 else 128 if K > 64 else 64 if K > 32 else 32)) if tc else 'direct',
             'flops': 2.0 * N * out_hw[0] * out_hw[1] * w.shape[1] * w.shape[2] * C * K,
             'bytes': float(esz * (N * H * W * C + w.numel()) + y.element_size() * N * out_hw[0] * out_hw[1] * K +
                            (esz * N * out_hw[0] * out_hw[1] * K if residual is not None else 0)),
             'sig': (N, H, W, C, K, w.shape[1], stride, dilation, residual is not None, bn_sum is not None or bnb is not None,
                     'fprop'),
             'e0': torch.cuda.Event(enable_timing=True), 'e1': torch.cuda.Event(enable_timing=True)}
      rec['e0'].record()
    if bn_sum is not None and not tc:
      # direct kernel: statistics from the stored output instead of the accumulators
      ops.conv2d_fprop(prm, x, w, y, scale, shift, residual)
      ops.bn_stats(y, N * out_hw[0] * out_hw[1], K, y.stride(2), bn_sum, bn_sqsum)
      if fin is not None:
        ops.bn_finalize(bn_sum, bn_sqsum, fin[0], K, *fin[1:-1])
    elif fin is not None:
      # statistics + finalisation in the convolution launch (fin = the arguments of ops.bn_finalize after the sums)
      assert scale is None and residual is None and not relu
      ops.conv2d_fprop_bn(prm, x, w, y, bn_sum, bn_sqsum, *fin)
    elif bnb is not None:
      ops.conv2d_fprop_bnbwd(prm, x, w, y, *bnb)
    elif out_mask is not None:
      assert scale is None and bn_sum is None and not relu
      ops.conv2d_fprop_masked(prm, x, w, y, residual, out_mask)
    else:
      ops.conv2d_fprop(prm, x, w, y, scale, shift, residual, bn_sum, bn_sqsum)
    if rec is not None:
      rec['e1'].record()
      self.profile.append(rec)
    return y

  # ---- inference ---------------------------------------------------------------------------------
  def _conv_bn_infer(self, x, scope, relu=None, residual=None, res_stride=1):
    spec = self.p.by_scope[scope]
    pt, pl, P, Q = self._geom(spec, x.shape[1], x.shape[2])
    scale, shift = self.p.folded_bn(scope, eps=self.eps)
    return self._conv(x, self._weights(scope), stride=spec.stride, dilation=spec.dilation, pad=(pt, pl),
                      out_hw=(P, Q), scale=scale, shift=shift, relu=spec.relu if relu is None else relu,
                      residual=residual, res_stride=res_stride)

  def _root_infer(self, images):
    scope = f'{arch.RES}/conv1'
    N, H, W, _ = images.shape
    scale, shift = self.p.folded_bn(scope, eps=self.eps)
    if self.dtype == torch.bfloat16:
      Hs, Ws = (H + 1) // 2, (W + 1) // 2
      packed = torch.empty((N, Hs, Ws, 64), dtype=torch.bfloat16, device=self.dev)
      ops.conv1_pack(images, packed)
      w2 = self.p.conv1_packed_weights(torch.bfloat16)
      y = self._conv(packed, w2, pad=(2, 0), out_hw=(Hs, Ws), scale=scale, shift=shift, relu=True)
    else:
      y = self._conv_bn_infer(images.to(self.dtype), scope)
    P, Q = (y.shape[1] + 1) // 2, (y.shape[2] + 1) // 2
    pooled = torch.empty((N, P, Q, 64), dtype=self.dtype, device=self.dev)
    ops.maxpool_same_fwd(y, pooled, 3, 2)
    return pooled

  def _unit_infer(self, x, u):
    if u.has_shortcut_conv:
      shortcut, rs = self._conv_bn_infer(x, f'{u.scope}/shortcut'), 1
    else:
      # identity shortcut; a strided unit subsamples it (1x1 max-pool, stride s) -> strided read
      shortcut, rs = x, u.stride
    r = self._conv_bn_infer(x, f'{u.scope}/conv1')
    r = self._conv_bn_infer(r, f'{u.scope}/conv2')
    return self._conv_bn_infer(r, f'{u.scope}/conv3', relu=True, residual=shortcut, res_stride=rs)

  def features_infer(self, images):
    x = self._root_infer(images)
    for u in arch.units():
      x = self._unit_infer(x, u)
    f = self._conv_bn_infer(x, 'feature_extractor/extension/decrease_fdims')
    if self.p.fov:
      f = self._conv_bn_infer(f, arch.FOV_SCOPE)
    return self._psp(f, self._conv_bn_infer) if self.p.psp else f

  def _psp(self, bottom, layer):
    """_create_psp_module (models/resnet50_extended_model_hierarchical.py:186-207): VALID average pooling
    into 1 / 2 / 3 / 6 bins -> 1x1 conv (+BN+ReLU) -> bilinear (align_corners) back to h x w, written
    straight into its channel slice of the [N, h, w, 5d] concatenation -> 1x1 conv.  `layer(x, scope)` is
    the conv+BN(+ReLU) driver of the calling mode (inference or training)."""
    N, h, w, d = bottom.shape
    cat = torch.empty((N, h, w, 5 * d), dtype=bottom.dtype, device=self.dev)
    cat[..., :d].copy_(bottom)
    pooled_shapes = []
    for i, (sc, b) in enumerate(zip(arch.PSP_SCOPES[:4], arch.PSP_BINS)):
      kh, kw = h // b, w // b
      assert kh > 0 and kw > 0, f'--psp_module: the {h}x{w} feature map is too small for {b} bins'
      pooled = torch.empty((N, (h - kh) // kh + 1, (w - kw) // kw + 1, d), dtype=bottom.dtype, device=self.dev)
      ops.avgpool_valid_fwd(bottom, pooled, kh, kw)
      c = layer(pooled, sc)
      ops.resize_bilinear_fwd(c, cat[..., (i + 1) * d:(i + 2) * d])
      pooled_shapes.append((kh, kw, tuple(pooled.shape)))
    self._psp_geom = pooled_shapes
    return layer(cat, arch.PSP_SCOPES[4])

  def lowres_logits_infer(self, images):
    """fp32 [N, h, w, logits_pitch]: the three heads' logits, concatenated along channels."""
    f = self.features_infer(images)
    N, h, w, d = f.shape
    au = arch.adaptation_units(d)
    # the three adaptation conv1 kernels are adjacent in the arena: one 256 -> 768 GEMM
    s0 = f'{au[0].scope}/conv1'
    o = self.p.w_off[s0]
    arena = self.p.operand if self.dtype == torch.bfloat16 else self.p.master
    w1 = arena[o:o + 3 * d * d].view(3 * d, 1, 1, d)
    scale, shift = self.p.folded_bn(s0, n=3 * d, eps=self.eps)
    a1 = self._conv(f, w1, out_hw=(h, w), scale=scale, shift=shift, relu=True)
    pitch = self.hier.logits_pitch
    logits = torch.zeros((N, h, w, pitch), dtype=torch.float32, device=self.dev)
    c0 = 0
    for b, (u, (_, lg)) in enumerate(zip(au, arch.BRANCHES)):
      r = self._conv_bn_infer(a1[..., b * d:(b + 1) * d], f'{u.scope}/conv2')
      r = self._conv_bn_infer(r, f'{u.scope}/conv3', relu=True, residual=f)
      scope = f'softmax_classifier/{lg}'
      ck = self.p.by_scope[scope].K
      sc, sh = self.p.folded_bn(scope, eps=self.eps)
      self._conv(r, self._weights(scope), out_hw=(h, w), scale=sc, shift=sh, relu=False,
                 y=logits[..., c0:c0 + ck])
      c0 += ck
    return self._hybrid_upsampler_fwd(logits) if self.p.upsampling == 'hybrid' else logits

  def _hybrid_upsampler_fwd(self, logits):
    """`--upsampling_method hybrid` (models/resnet50_extended_model_hierarchical.py:168-180): per head a 3x3
    slim.conv2d_transpose (stride 1, SAME, + bias) on the low-resolution logits, before the bilinear resize the
    head / loss kernels do.  fp32 on the direct kernel (14 / 7 / 3 channels), reading and writing channel slices
    of the pitched logits buffers; the transposed convolution is run as the stride-1 correlation it equals."""
    N, h, w, pitch = logits.shape
    out = torch.zeros_like(logits)
    ones = torch.ones(64, dtype=torch.float32, device=self.dev)
    c0 = 0
    for sc in arch.UPSAMPLING_SCOPES:
      spec = self.p.by_scope[sc]
      ck = spec.K
      prm = ops.conv_params((N, h, w, ck), (ck, 3, 3, ck), pad=(1, 1), out_hw=(h, w), x_pitch=pitch, y_pitch=pitch,
                            dtype=ops.F32, algo=ops.ALGO_DIRECT)
      ops.conv2d_fprop(prm, logits[..., c0:c0 + ck], self.p.w32(sc), out[..., c0:c0 + ck], ones[:ck], self.p.beta(sc))
      c0 += ck
    return out

  def predict(self, images, want=('decisions',)):
    """Forward pass -> dict with the requested keys of the reference's predictions dict
    (code/models/resnet50_extended_model_hierarchical.py:121-130)."""
    N, H, W, _ = images.shape
    logits = self.lowres_logits_infer(images)
    if self.p.upsampling == 'no':   # `upsampled = bottom` (:164-165): predictions at the feature resolution
      H, W = logits.shape[1], logits.shape[2]
    out = {'lowres_logits': logits}
    out.update(self.head(logits, H, W, want))
    return out

  def head(self, logits, H, W, want=('decisions',)):
    N = logits.shape[0]
    C1, Cv, Ch = self.hier.head_widths
    dev = self.dev

    def i32(key):
      return torch.empty((N, H, W), dtype=torch.int32, device=dev) if key in want else None

    def f32(key, c):
      return torch.empty((N, H, W, c), dtype=torch.float32, device=dev) if key in want else None
    res = {'decisions': i32('decisions'), 'l1_decisions': i32('l1_decisions'),
           'l2_vehicle_decisions': i32('l2_vehicle_decisions'), 'l2_human_decisions': i32('l2_human_decisions'),
           'l1_probabilities': f32('l1_probabilities', C1), 'l2_vehicle_probabilities': f32('l2_vehicle_probabilities', Cv),
           'l2_human_probabilities': f32('l2_human_probabilities', Ch)}
    full = None
    if any(k in want for k in ('l1_logits', 'l2_vehicle_logits', 'l2_human_logits', 'logits')):
      full = torch.empty((N, H, W, C1 + Cv + Ch), dtype=torch.float32, device=dev)
    # algorithmic HBM bytes: every requested full-resolution map written once + the low-res logits
    out_bytes = sum(v.numel() * v.element_size() for v in res.values() if v is not None)
    out_bytes += 0 if full is None else full.numel() * 4
    with self._Timed(self, 'head_fwd', out_bytes + N * logits.shape[1] * logits.shape[2] * (C1 + Cv + Ch) * 4):
      ops.head_fwd(self.hstruct, logits, H, W, res['decisions'], res['l1_decisions'], res['l2_vehicle_decisions'],
                   res['l2_human_decisions'], res['l1_probabilities'], res['l2_vehicle_probabilities'],
                   res['l2_human_probabilities'], full)
    out = {k: v for k, v in res.items() if v is not None}
    if full is not None:
      out['l1_logits'] = full[..., :C1]
      out['l2_vehicle_logits'] = full[..., C1:C1 + Cv]
      out['l2_human_logits'] = full[..., C1 + Cv:]
    return out


class EvalStep:
  """One evaluation step - forward, hierarchical decisions, (optional void replacement,) confusion-matrix update -
  replayed as ONE CUDA graph (define_estimator EVAL branch, code/estimator/define_estimator_hierarchical.py:161-202).

  The eagerly launched step leaves the GPU idle for ~0.7 of its 8.9 ms at 4 x 1024 x 2048 (69 launches from Python,
  profiles/r1_timeline_eval.txt).  A graph is keyed by the ADDRESSES of its inputs: an input pipeline that rotates
  through a few device buffers (the bench's two resident batches, a double-buffered loader) replays without any
  copy; inputs at new addresses are copied into one static pair first.  Labels and network must agree in size
  (the resize of `_resize_predictions` is an identity then); other cases take the eager path of the caller."""

  MAX_POINTER_GRAPHS = 4

  def __init__(self, net, num_classes, lut=None, replace_voids=False, void_cid=None):
    self.net, self.num_classes, self.lut = net, num_classes, lut
    self.replace_voids, self.void_cid = replace_voids, void_cid
    dev = net.dev
    self.cm = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=dev)
    self.invalid = torch.zeros(1, dtype=torch.int64, device=dev)
    self._graphs = {}
    self._static = {}
    self._warm = set()
    self._version = net.p.version
    self.enabled = dev.type == 'cuda'

  def reset(self):
    self.cm.zero_()
    self.invalid.zero_()

  def _body(self, images, labels):
    want = ('decisions',) + (('l1_probabilities', 'l2_vehicle_probabilities', 'l2_human_probabilities')
                             if self.replace_voids else ())
    if not self.replace_voids and os.environ.get('WLSEG_HEAD_CONFMAT', '1') != '0':
      # the step's tail in one launch: decisions and labels meet in registers (ops.head_confmat)
      low = self.net.lowres_logits_infer(images)
      H, W = labels.shape[1], labels.shape[2]
      N = low.shape[0]
      nbytes = N * H * W * 4 + N * low.shape[1] * low.shape[2] * sum(self.net.hier.head_widths) * 4
      with self.net._Timed(self.net, 'head_confmat', nbytes):
        ops.head_confmat(self.net.hstruct, low, H, W, labels, self.num_classes, self.cm, self.lut, self.invalid)
      return
    out = self.net.predict(images, want=want)
    if self.replace_voids:
      ops.replace_voids(self.net.hstruct, out['l1_probabilities'], out['l2_vehicle_probabilities'],
                        out['l2_human_probabilities'], out['decisions'], self.void_cid)
    ops.confmat_accumulate(labels, out['decisions'], self.num_classes, self.cm, self.lut, self.invalid)

  def _capture(self, images, labels):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      self._body(images, labels)
    return g

  def __call__(self, images, labels):
    shape = (tuple(images.shape), images.dtype, tuple(labels.shape))
    if (not self.enabled or self.net.profile is not None or tuple(labels.shape[1:3]) != tuple(images.shape[1:3])
            or not labels.is_contiguous() or not images.is_contiguous()):
      return self._body(images, labels)   # callers with per-launch events, odd layouts: eager
    if self._version != self.net.p.version:
      # the parameters changed: operands derived from them (folded BN constants, packed root filters) were
      # re-created at new addresses, which the captured graphs do not know
      self._graphs.clear()
      self._static.clear()
      self._warm.clear()
      self._version = self.net.p.version
    if shape not in self._warm:           # first step of a shape runs eagerly (lazy kernel attributes, allocator)
      self._warm.add(shape)
      return self._body(images, labels)
    key = (images.data_ptr(), labels.data_ptr()) + shape
    g = self._graphs.get(key)
    if g is None and sum(1 for k in self._graphs if k[2:] == shape) < self.MAX_POINTER_GRAPHS:
      g = self._graphs[key] = (self._capture(images, labels), images, labels)   # keeps the buffers alive
    if g is not None:
      g[0].replay()
      return
    st = self._static.get(shape)
    if st is None:
      si, sl = torch.empty_like(images), torch.empty_like(labels)
      si.copy_(images)
      sl.copy_(labels)
      st = self._static[shape] = (self._capture(si, sl), si, sl)
    st[1].copy_(images, non_blocking=True)
    st[2].copy_(labels, non_blocking=True)
    st[0].replay()


# ================================================================================================
# Training: forward with batch statistics, fused loss forward/backward, backward pass
# ================================================================================================
class _Rec:
  """Tape entry of one conv + BN (+ residual) (+ ReLU) layer."""
  __slots__ = ('scope', 'spec', 'x', 'w', 'z', 'a', 'relu', 'has_res', 'geom', 'nch', 'kind', 'gn', 'mask',
               'res', 'da', 'dz', 'dx', 'dx_add')  # the last five only when TrainNetwork.keep (tests)


class TrainWorkspace:
  """Per-step scratch that lives across steps: BN statistic accumulators (fp64), per-layer BN
  scale / shift / saved mean / inverse std (fp32), the gradient arena (same layout as the master
  parameter arena) and the momentum / EMA arenas."""

  def __init__(self, params):
    p = params
    dev = p.device
    n = p.n_chan_pad
    self.stat = torch.zeros(4 * n, dtype=torch.float64, device=dev)  # [sum | sqsum | dgamma | dbeta]
    self.bn = torch.zeros(4 * n, dtype=torch.float32, device=dev)    # [scale | shift | mean | invstd]
    self.grads = torch.zeros(p.n_total, dtype=torch.float32, device=dev)
    self.momentum = torch.zeros(p.n_total, dtype=torch.float32, device=dev)
    self.ema_biased = None
    self.ema_shadow = None
    self.loss_sums = torch.zeros(3, dtype=torch.float64, device=dev)
    self.loss_counts = torch.zeros(3, dtype=torch.float64, device=dev)
    self.losses = torch.zeros(4, dtype=torch.float32, device=dev)
    self.reg_loss = torch.zeros(1, dtype=torch.float64, device=dev)
    self.lr = torch.zeros(1, dtype=torch.float32, device=dev)
    self.fin_counter = torch.zeros(1, dtype=torch.int32, device=dev)   # ops.conv2d_fprop_bn's ticket word (left at zero)
    self.n = n

  def view(self, arena, slot, off, k):
    return arena[slot * self.n + off:slot * self.n + off + k]


class TrainNetwork(Network):
  """Forward with batch statistics + backward.  BN follows every convolution (also the logits
  convolutions, code/models/resnet50_extended_model_hierarchical.py:76-83); statistics are per
  replica (the reference's default) or, with `cross_replica=(world_size, process_group)`
  (--cross_replica_norm, code/utils/cross_replica_batch_normalization.py:398-459), over all replicas:
  the per-layer fp64 [sum | sqsum] (forward) and [dgamma | dbeta] (backward) vectors are summed
  across ranks with one small all-reduce each, between the two passes that produce and consume them."""

  def __init__(self, params, dtype=torch.bfloat16, bn_decay=0.9, eps=1e-5, conv_algo=ops.ALGO_AUTO,
               cross_replica=None):
    super().__init__(params, dtype, bn_decay, eps, conv_algo)
    self.ws = TrainWorkspace(params)
    self.cross_replica = cross_replica if (cross_replica and cross_replica[0] >= 1) else None  # (1, group): tests
    self.grad_ready = None  # callback(lo): every conv-kernel gradient at arena offset >= lo is final
    self.keep = False       # tests: keep per-layer gradient tensors on the tape
    self._flipped = None    # {scope: dgrad filter bank view}, refreshed at the start of backward()
    # one launch for bn_finalize + bn_apply: OFF.  Round 1's form (every CTA staged all C constants in shared memory)
    # measured 0.15 ms/step slower than two launches inside the step graph; round 2's form (csrc/bn.cu BnFin: a thread
    # derives only its own 8 channels, fp64 multiplications instead of divisions) still loses, 352 against 361
    # images/s in two A/B pairs on one box: the dependent chain fp64 loads -> rsqrt -> constants in front of every
    # thread's first streaming load costs more than a 3 us launch of the finalisation kernel.  WLSEG_BN_FUSED_FINALIZE=1
    self.fused_bn_finalize = os.environ.get('WLSEG_BN_FUSED_FINALIZE', '0') == '1'
    # per-slice working set of the two BN backward passes.  Measured on B200 (profiles/): 64 MB slices
    # (second pass from L2) LOSE to whole-tensor passes - 512-byte row segments at a 4 KB pitch halve
    # the HBM efficiency of the first pass - so slicing is off by default.
    self.bn_bwd_l2_bytes = 1 << 40
    self._order = {s.scope: i for i, s in enumerate(params.specs)}
    # filter gradients on a second stream: wgrad of layer L only needs dz_L, while the chain dgrad_L ->
    # BN backward of layer L-1 -> ... does not need it, so the (tensor-core) wgrad kernels can fill the SMs
    # next to the (bandwidth) BN backward kernels of the layers below.  None = everything on one stream.
    self.wgrad_stream = None
    self._side_keep = []
    # ReLU bit masks of the bottleneck outputs (1 bit per element, written by their bn_apply): the dgrad that
    # completes the gradient of such a tensor multiplies it by the mask in its epilogue, so the unit's BN backward
    # neither reads the activation nor writes a separate shortcut gradient (it IS the masked tensor): 3 of the 8
    # tensor passes of every residual layer's BN backward.  WLSEG_RELU_MASK=0 restores the unmasked wiring.
    self.premask = os.environ.get('WLSEG_RELU_MASK', '1') != '0'
    # BN backward reduction fused into the dgrad that produces the gradient it reduces (ops.conv2d_fprop_bnbwd): inside a
    # bottleneck unit the dgrad of conv3 / conv2 multiplies its output by the ReLU derivative of conv2 / conv1's BN and
    # accumulates that layer's dgamma / dbeta, so 32 of the 61 bn_bwd_reduce passes of a step are never launched.
    # WLSEG_BNB_FUSE=0 restores the separate passes.
    self.bnb_fuse = os.environ.get('WLSEG_BNB_FUSE', '1') != '0'
    # bn_finalize inside the convolution launch that produces the statistics (ops.conv2d_fprop_bn: the last CTA to
    # commit its sums finalises): 61 launches of ~3 us less per step - and a measured NEGATIVE, 11.00 -> 11.18 ms/step
    # (357.7 against 363.6 images/s, A/B on one box): every CTA pays a __threadfence behind its fp64 atomics plus two
    # barriers, and the finalisation's dependent fp64 chain (L2 loads -> divisions -> rsqrt) runs on ONE SM behind
    # the grid's tail instead of next to it; the convolution launches grew by more (+0.33 ms) than the 0.21 ms the
    # bn_finalize launches cost.  Kept as an opt-in: WLSEG_BN_FIN_IN_CONV=1.
    self.fin_in_conv = os.environ.get('WLSEG_BN_FIN_IN_CONV', '0') == '1'

  def _mark_done(self, scope, n_elems):
    """Bookkeeping for the gradient exchange: backward completes the arena roughly tail first."""
    if self.grad_ready is None:
      return
    lo = self.p.w_off[scope]
    i = self._order[scope]
    specs = self.p.specs
    while i < len(specs) and self.p.w_off[specs[i].scope] < lo + n_elems:
      self._done[i] = True
      i += 1
    t = self._tail
    while t > 0 and self._done[t - 1]:
      t -= 1
    if t != self._tail:
      self._tail = t
      self.grad_ready(self.p.w_off[specs[t].scope])

  # ---- one layer ---------------------------------------------------------------------------------
  def _layer_fwd(self, x, scope, *, w=None, nch=None, relu=None, residual=None, pad=None, stride=None,
                 dilation=None, out_hw=None, y_f32=False, kind='conv', want_mask=False):
    spec = self.p.by_scope[scope]
    if w is None:
      w = self._weights(scope)
    K = w.shape[0] if nch is None else nch
    if pad is None:
      pt, pl, P, Q = self._geom(spec, x.shape[1], x.shape[2])
      pad, out_hw, stride, dilation = (pt, pl), (P, Q), spec.stride, spec.dilation
    N = x.shape[0]
    count = N * out_hw[0] * out_hw[1]
    ws, off = self.ws, self.p.c_off[scope]
    s1, s2 = ws.view(ws.stat, 0, off, K), ws.view(ws.stat, 1, off, K)
    zdt = torch.float32 if y_f32 else self.dtype
    group = self.p.norm == 'group'
    scale, shift = ws.view(ws.bn, 0, off, K), ws.view(ws.bn, 1, off, K)
    mean, invstd = ws.view(ws.bn, 2, off, K), ws.view(ws.bn, 3, off, K)
    fin = None
    if (self.fin_in_conv and not group and self.cross_replica is None and not self.fused_bn_finalize and not y_f32 and
        self.dtype == torch.bfloat16 and self.conv_algo != ops.ALGO_DIRECT and K % 64 == 0 and x.shape[3] % 8 == 0 and
        x.stride(2) % 8 == 0 and not ops.pdl_enabled()):
      fin = (count, self.p.gamma(scope, K), self.p.beta(scope, K), self.eps, self.bn_decay, self.p.moving_mean(scope, K),
             self.p.moving_var(scope, K), scale, shift, mean, invstd, ws.fin_counter)
    z = self._conv(x, w, stride=stride, dilation=dilation, pad=pad, out_hw=out_hw, y_dtype=zdt,
                   bn_sum=None if group else s1, bn_sqsum=None if group else s2, fin=fin)
    a = torch.empty_like(z)
    do_relu = spec.relu if relu is None else relu
    gn = None
    mask = None
    if group:
      gn = self._gn_forward(z, scope, K, do_relu, residual, a)
    elif self.cross_replica is not None:
      # global moments: one all-reduce of this layer's [sum | sqsum]; every replica has the same count
      R = self.cross_replica[0]
      both = self._all_reduce_pair(s1, s2)
      ops.bn_finalize(both[:K], both[K:], count * R, K, self.p.gamma(scope, K), self.p.beta(scope, K), self.eps,
                      self.bn_decay, self.p.moving_mean(scope, K), self.p.moving_var(scope, K), scale, shift, mean,
                      invstd, moving_var_factor=(count - 1.0) / count)
      ops.bn_apply(z, scale, shift, residual, a, count, K, do_relu)
    elif self.fused_bn_finalize and K % 8 == 0 and K <= 2048 and z.is_contiguous():
      if want_mask and do_relu and K % 32 == 0:
        mask = torch.empty((count, K // 8), dtype=torch.uint8, device=self.dev)
      ops.bn_finalize_apply(s1, s2, count, K, self.p.gamma(scope, K), self.p.beta(scope, K), self.eps, self.bn_decay,
                            self.p.moving_mean(scope, K), self.p.moving_var(scope, K), scale, shift, mean, invstd,
                            z, residual, a, do_relu, mask=mask)
    else:
      if fin is None:
        ops.bn_finalize(s1, s2, count, K, self.p.gamma(scope, K), self.p.beta(scope, K), self.eps, self.bn_decay,
                        self.p.moving_mean(scope, K), self.p.moving_var(scope, K), scale, shift, mean, invstd)
      if want_mask and do_relu and K % 32 == 0 and K <= 2048 and z.is_contiguous():
        mask = torch.empty((count, K // 8), dtype=torch.uint8, device=self.dev)
      ops.bn_apply(z, scale, shift, residual, a, count, K, do_relu, mask=mask)
    rec = _Rec()
    rec.scope, rec.spec, rec.x, rec.w, rec.z, rec.a = scope, spec, x, w, z, a
    rec.relu, rec.has_res, rec.geom, rec.nch, rec.kind = do_relu, residual is not None, (pad, out_hw, stride, dilation), K, kind
    rec.gn = gn
    rec.mask = mask
    rec.res = residual if self.keep else None
    self.tape[scope] = rec
    return a

  # ---- group norm (--norm_layer group): the per-sample passes are the batch-norm kernels on one sample's rows ----
  def _gn_forward(self, z, scope, K, relu, residual, a):
    """z [N, P, Q, K] -> a = relu?(group_norm(z) (+ residual)); returns the per-(sample, channel) rows
    [scale | shift | mean | invstd] (fp32 [4, N, K]) the backward pass needs."""
    N, P, Q, _ = z.shape
    hw = P * Q
    sums = torch.zeros((2, N, K), dtype=torch.float64, device=self.dev)
    for n in range(N):
      ops.bn_stats(z[n], hw, K, z.stride(2), sums[0, n], sums[1, n])
    coef = torch.empty((4, N, K), dtype=torch.float32, device=self.dev)
    ops.gn_finalize(sums[0], sums[1], N, K, self.p.groups(scope, K), hw, self.p.gamma(scope, K), self.p.beta(scope, K),
                    self.eps, coef[0], coef[1], coef[2], coef[3])
    for n in range(N):
      ops.bn_apply(z[n], coef[0, n], coef[1, n], None if residual is None else residual[n], a[n], hw, K, relu)
    return coef

  def _gn_backward(self, rec, da, scope, K, dgamma, dbeta):
    """-> (dz, dres) of a group-normalised layer; dgamma / dbeta (fp64 accumulators) receive the parameter sums."""
    N, P, Q, _ = rec.z.shape
    hw = P * Q
    coef = rec.gn
    y = rec.a if (rec.relu and (rec.has_res or K % 8 != 0)) else None
    da = da.contiguous()
    part = torch.zeros((2, N, K), dtype=torch.float64, device=self.dev)
    for n in range(N):
      ops.bn_bwd_reduce(da[n], None if y is None else y[n], rec.z[n], coef[2, n], coef[3, n], hw, K, rec.relu,
                        part[0, n], part[1, n], scale=coef[0, n], shift=coef[1, n], pitch=K)
    k = torch.empty((3, N, K), dtype=torch.float32, device=self.dev)
    ops.gn_bwd_finalize(part[0], part[1], N, K, self.p.groups(scope, K), hw, self.p.gamma(scope, K), coef[2], coef[3],
                        k[0], k[1], k[2], dgamma, dbeta)
    dz = torch.empty_like(rec.z)
    dres = torch.empty_like(rec.z) if rec.has_res else None
    ops.gn_bwd_apply(da, y, rec.z, k[0], k[1], k[2], coef[0], coef[1], N, hw, K, rec.relu, dz, dres)
    return dz, dres

  def lowres_logits_infer(self, images):
    """Group norm has no inference form of its own (no moving statistics): the evaluation forward IS the
    training forward.  Batch norm keeps the folded-scale/shift path of the base class."""
    if self.p.norm == 'group':
      return self.forward_train(images)
    return super().lowres_logits_infer(images)

  def _all_reduce_pair(self, a, b):
    """SUM over replicas of two per-channel fp64 vectors with ONE collective; returns [a | b] summed."""
    import torch.distributed as dist
    both = torch.cat([a, b])
    dist.all_reduce(both, op=dist.ReduceOp.SUM, group=self.cross_replica[1])
    return both

  def _wgrad_view(self, scope, shape):
    o = self.p.w_off[scope]
    n = 1
    for d in shape:
      n *= d
    return self.ws.grads[o:o + n].view(*shape)

  def _bnb_args(self, scope):
    """-> the bnb tuple of `scope` (the layer whose OUTPUT activation a dgrad differentiates), or None when the fused
    form does not cover it."""
    if not (self.bnb_fuse and not self.keep and self.dtype == torch.bfloat16 and self.conv_algo != ops.ALGO_DIRECT and
            self.p.norm == 'batch' and self.cross_replica is None):
      return None
    rec = self.tape[scope]
    K = rec.nch
    if not (rec.relu and not rec.has_res and K % 64 == 0 and rec.z.is_contiguous() and rec.z.dtype == torch.bfloat16):
      return None
    ws, off = self.ws, self.p.c_off[scope]
    if off % 4 != 0:
      return None   # scale / shift are read as float4
    return (rec.z, ws.view(ws.bn, 0, off, K), ws.view(ws.bn, 1, off, K), ws.view(ws.bn, 2, off, K),
            ws.view(ws.bn, 3, off, K), ws.view(ws.stat, 2, off, K), ws.view(ws.stat, 3, off, K))

  def _layer_bwd(self, scope, da, need_dx=True, dx_out=None, dx_add=None, da_masked=False, dx_mask=None,
                 da_reduced=False, bnb_below=None):
    """Backward of one tape entry.  Returns (dx, dres): dres is the gradient of the residual input
    (None if the layer had none).  `dx_add` is accumulated into dx (fused into the tensor-core
    dgrad epilogue when possible); `dx_out` lets the caller place dx (channel-sliced views).
    da_masked: `da` already carries this layer's ReLU derivative (its producer applied rec.mask): the BN backward
    reads da and z only and dres is da itself.  dx_mask: ReLU bit mask of the tensor dx belongs to, applied by
    the dgrad epilogue (dx + dx_add is that tensor's COMPLETE gradient).
    da_reduced: the dgrad that produced `da` ran in the BN-backward form - da carries this layer's ReLU derivative and
    dgamma / dbeta are already accumulated: only bn_bwd_apply runs.  bnb_below: scope of the layer that produced this
    layer's INPUT; this layer's dgrad then runs in that form for it (the caller passes da_reduced to it)."""
    rec = self.tape[scope]
    ws, off, K = self.ws, self.p.c_off[scope], rec.nch
    pad, out_hw, stride, dilation = rec.geom
    N, H, W, C = rec.x.shape
    count = N * out_hw[0] * out_hw[1]
    mean, invstd = ws.view(ws.bn, 2, off, K), ws.view(ws.bn, 3, off, K)
    dgamma, dbeta = ws.view(ws.stat, 2, off, K), ws.view(ws.stat, 3, off, K)
    scale, shift = ws.view(ws.bn, 0, off, K), ws.view(ws.bn, 1, off, K)
    gamma = self.p.gamma(scope, K)
    if self.p.norm == 'group':
      assert not da_masked and dx_mask is None
      dz, dres = self._gn_backward(rec, da, scope, K, dgamma, dbeta)
      return self._layer_bwd_convs(rec, scope, da, dz, dres, K, need_dx, dx_out, dx_add)
    dz = torch.empty_like(rec.z)
    relu = rec.relu
    if da_reduced:
      assert not rec.has_res and rec.relu and da.is_contiguous() and not da_masked
      dres, relu = None, False
      ops.bn_bwd_apply(da, None, rec.z, mean, invstd, gamma, dgamma, dbeta, count, K, False, dz, None, scale=scale,
                       shift=shift, pitch=K)
      return self._layer_bwd_convs(rec, scope, da, dz, None, K, need_dx, dx_out, dx_add, dx_mask, bnb_below)
    if da_masked:
      assert rec.has_res and rec.relu and da.is_contiguous()
      dres, relu = da, False          # g = da: nothing left to mask, and it is the shortcut's gradient as it stands
    else:
      dres = torch.empty_like(rec.z) if rec.has_res else None
    # the ReLU mask of a layer without a residual input is recomputed from z (y is not read at all)
    y = rec.a if (relu and (rec.has_res or K % 8 != 0)) else None
    # channel slices sized so that one slice of dy / y / z stays L2 resident between the two passes
    esz = rec.z.element_size()
    Kg = K
    if K % 8 == 0 and da.is_contiguous() and rec.z.is_contiguous():
      budget = self.bn_bwd_l2_bytes // ((3 if y is not None else 2) * esz * max(count, 1))
      if budget < K:
        Kg = max(64, budget // 64 * 64)
    if self.cross_replica is not None:
      # the dz formula needs the sums over ALL replicas' pixels (the all-reduce of the forward moments is its
      # own transpose); the parameter gradients stay the local sums and are averaged with the other gradients
      R = self.cross_replica[0]
      ops.bn_bwd_reduce(da, y, rec.z, mean, invstd, count, K, relu, dgamma, dbeta, scale=scale, shift=shift, pitch=K)
      both = self._all_reduce_pair(dgamma, dbeta)
      ops.bn_bwd_apply(da, y, rec.z, mean, invstd, gamma, both[:K], both[K:], count, K, relu, dz,
                       None if da_masked else dres, scale=scale, shift=shift, pitch=K, stat_count=count * R)
      Kg = 0
    for c0 in (range(0, K, Kg) if Kg else ()):
      kk = min(Kg, K - c0)
      sl = slice(c0, c0 + kk)
      ops.bn_bwd_reduce(da[..., sl], None if y is None else y[..., sl], rec.z[..., sl], mean[sl], invstd[sl], count, kk,
                        relu, dgamma[sl], dbeta[sl], scale=scale[sl], shift=shift[sl], pitch=K)
      ops.bn_bwd_apply(da[..., sl], None if y is None else y[..., sl], rec.z[..., sl], mean[sl], invstd[sl], gamma[sl],
                       dgamma[sl], dbeta[sl], count, kk, relu, dz[..., sl],
                       None if (dres is None or da_masked) else dres[..., sl], scale=scale[sl], shift=shift[sl], pitch=K)
    return self._layer_bwd_convs(rec, scope, da, dz, dres, K, need_dx, dx_out, dx_add, dx_mask, bnb_below)

  def _layer_bwd_convs(self, rec, scope, da, dz, dres, K, need_dx, dx_out, dx_add, dx_mask=None, bnb_below=None):
    """Second half of _layer_bwd: filter gradient and data gradient from dz (normaliser independent)."""
    pad, out_hw, stride, dilation = rec.geom
    N, H, W, C = rec.x.shape
    wsrc = rec.w
    if dz.dtype != self.dtype or (self.dtype == torch.bfloat16 and K % 8 != 0):
      # logits layers (fp32 z, 14/7/3 channels) feeding bf16 tensor-core kernels: bf16 copy of dz with
      # the channel count padded to a multiple of 8 (zero columns), TMA needs 16-byte pixel pitches
      kp = (K + 7) // 8 * 8 if self.dtype == torch.bfloat16 else K
      dzc = torch.zeros(dz.shape[:-1] + (kp,), dtype=self.dtype, device=self.dev)
      dzc[..., :K] = dz
      dz = dzc
    # ---- filter gradient
    if rec.kind == 'root_packed':
      self._root_wgrad(rec, dz)
    else:
      prm = ops.conv_params((N, H, W, C), (K,) + tuple(rec.w.shape[1:]), stride=stride, dilation=dilation, pad=pad,
                            out_hw=out_hw, x_pitch=rec.x.stride(2), y_pitch=dz.stride(2), dtype=self.code,
                            accumulate=True)  # the gradient arena was zeroed once, at the start of backward()
      prof = None
      if self.profile is not None:
        Rr, Ss = rec.w.shape[1], rec.w.shape[2]
        prof = {'cls': 'wgrad_tc' if (self.dtype == torch.bfloat16 and C % 8 == 0) else 'wgrad_direct',
                'flops': 2.0 * N * out_hw[0] * out_hw[1] * Rr * Ss * C * K,
                'bytes': float(2 * (N * H * W * C + N * out_hw[0] * out_hw[1] * K) + 4 * K * Rr * Ss * C),
                'sig': (N, H, W, C, K, Rr, stride, dilation, False, False, 'wgrad'),
                'e0': torch.cuda.Event(enable_timing=True), 'e1': torch.cuda.Event(enable_timing=True)}
        prof['e0'].record()
      side = self.wgrad_stream if (prof is None and not self.keep) else None
      if side is None:
        ops.conv2d_wgrad(prm, rec.x, dz, self._wgrad_view(scope, (K,) + tuple(rec.w.shape[1:])))
      else:
        side.wait_event(torch.cuda.current_stream().record_event())   # dz (and the zeroed arena) are ready
        with torch.cuda.stream(side):
          ops.conv2d_wgrad(prm, rec.x, dz, self._wgrad_view(scope, (K,) + tuple(rec.w.shape[1:])))
        self._side_keep.append(dz)   # the allocator must not hand dz out again before the join in backward()
      if prof is not None:
        prof['e1'].record()
        self.profile.append(prof)
    self._mark_done(scope, K * rec.spec.R * rec.spec.S * rec.spec.C)
    if self.keep:
      rec.da, rec.dz, rec.dx, rec.dx_add = da, dz, None, dx_add
    if not need_dx:
      return None, dres
    # ---- data gradient
    dx = dx_out if dx_out is not None else torch.empty((N, H, W, C), dtype=self.dtype, device=self.dev)
    R, S = rec.w.shape[1], rec.w.shape[2]
    done = False
    if self.dtype == torch.bfloat16 and self.conv_algo != ops.ALGO_DIRECT:
      # dgrad == stride-1 fprop over dz (zero-inserted when the layer was strided) with the 180-degree
      # rotated, transposed filter bank
      Kp = dz.shape[-1]
      wf = self._flipped.get(scope) if self._flipped is not None else None
      if wf is None or tuple(wf.shape) != (C, R, S, K):
        wf = torch.empty((C, R, S, K), dtype=self.dtype, device=self.dev)
        ops.weights_transpose_flip(wsrc, wf)
      if Kp != K:
        wfp = torch.zeros((C, R, S, Kp), dtype=self.dtype, device=self.dev)
        wfp[..., :K] = wf
        wf = wfp
      if stride > 1:
        dzu = torch.empty((N, (out_hw[0] - 1) * stride + 1, (out_hw[1] - 1) * stride + 1, Kp), dtype=self.dtype,
                          device=self.dev)
        ops.zero_insert(dz, dzu, stride)
      else:
        dzu = dz
      fpad = (dilation * (R - 1) - pad[0], dilation * (S - 1) - pad[1])
      bnb = None
      if bnb_below is not None:
        assert dx_add is None and dx_mask is None and dx_out is None
        bnb = self._bnb_args(bnb_below)
        assert bnb is not None, 'the caller checks _bnb_args first'
      self._conv(dzu, wf, dilation=dilation, pad=fpad, out_hw=(H, W), residual=dx_add, y=dx, out_mask=dx_mask, bnb=bnb)
      done = True
    if not done:
      assert dx_mask is None, 'the masked data gradient exists on the tensor-core path only'
      prm = ops.conv_params((N, H, W, C), tuple(rec.w.shape), stride=stride, dilation=dilation, pad=pad,
                            out_hw=out_hw, x_pitch=dx.stride(2), y_pitch=dz.stride(2), dtype=self.code)
      ops.conv2d_dgrad(prm, dz, rec.w, dx)
      if dx_add is not None:
        assert dx.is_contiguous()
        ops.add_inplace(dx, dx_add)
    if self.keep:
      rec.dx = dx
    return dx, dres

  # ---- root ----------------------------------------------------------------------------------------
  def _root_fwd(self, images):
    scope = f'{arch.RES}/conv1'
    N, H, W, _ = images.shape
    if self.dtype == torch.bfloat16:
      Hs, Ws = (H + 1) // 2, (W + 1) // 2
      packed = torch.empty((N, Hs, Ws, 64), dtype=torch.bfloat16, device=self.dev)
      ops.conv1_pack(images, packed)
      self._root_images = images
      y = self._layer_fwd(packed, scope, w=self.p.conv1_packed_weights(torch.bfloat16), pad=(2, 0), stride=1,
                          dilation=1, out_hw=(Hs, Ws), relu=True, kind='root_packed')
    else:
      y = self._layer_fwd(images.to(self.dtype), scope)
    P, Q = (y.shape[1] + 1) // 2, (y.shape[2] + 1) // 2
    pooled = torch.empty((N, P, Q, 64), dtype=self.dtype, device=self.dev)
    amax = torch.empty((N, P, Q, 64), dtype=torch.uint8, device=self.dev)
    ops.maxpool_same_fwd(y, pooled, 3, 2, argmax=amax)
    self.tape['pool1'] = (y, amax)
    return pooled

  def _root_wgrad(self, rec, dz):
    """Filter gradient of the 7x7/2 root convolution, taken in the PACKED domain the forward GEMM
    runs in (R=4, S=1, C=64 over the space-to-depth tensor, on the tensor cores) and gathered back:
    every w[k, r, s, c] appears exactly once in the packed bank (conv1_packed_weights)."""
    scope = rec.scope
    N, Hs, Ws, _ = rec.x.shape
    pad, out_hw, stride, dilation = rec.geom
    dw2 = torch.empty((64, 4, 1, 64), dtype=torch.float32, device=self.dev)
    prm = ops.conv_params((N, Hs, Ws, 64), (64, 4, 1, 64), stride=stride, dilation=dilation, pad=pad, out_hw=out_hw,
                          x_pitch=rec.x.stride(2), y_pitch=dz.stride(2), dtype=self.code)
    ops.conv2d_wgrad(prm, rec.x, dz, dw2)
    idx = self.p.conv1_pack_index()
    self._wgrad_view(scope, (64, 7, 7, 3)).copy_(dw2.view(64, 256)[:, idx].view(64, 7, 7, 3))

  def _root_bwd(self, dpool):
    y, amax = self.tape['pool1']
    dy = torch.empty_like(y)
    ops.maxpool_same_bwd(None, dpool, dy, 3, 2, argmax=amax)
    self._layer_bwd(f'{arch.RES}/conv1', dy, need_dx=False)

  # ---- bottleneck units ------------------------------------------------------------------------------
  def _unit_fwd(self, x, u):
    if u.has_shortcut_conv:
      sc = self._layer_fwd(x, f'{u.scope}/shortcut')
    elif u.stride > 1:
      N, H, W, C = x.shape
      sc = torch.empty((N, (H + u.stride - 1) // u.stride, (W + u.stride - 1) // u.stride, C), dtype=self.dtype,
                       device=self.dev)
      ops.maxpool_same_fwd(x, sc, 1, u.stride)
    else:
      sc = x
    r = self._layer_fwd(x, f'{u.scope}/conv1')
    r = self._layer_fwd(r, f'{u.scope}/conv2')
    self.tape[u.scope] = x
    return self._layer_fwd(r, f'{u.scope}/conv3', relu=True, residual=sc, want_mask=self._use_premask())

  def _use_premask(self):
    return (self.premask and not self.keep and self.dtype == torch.bfloat16 and self.conv_algo != ops.ALGO_DIRECT and
            self.p.norm == 'batch')

  def _unit_bwd(self, dout, u, dout_masked=False, in_mask=None):
    """dout: gradient of the unit's output (already multiplied by its ReLU derivative if dout_masked);
    in_mask: ReLU bit mask of the unit's INPUT tensor (the previous unit's output) - the dgrad that completes dx
    applies it.  -> dx"""
    x = self.tape[u.scope]
    c1, c2 = f'{u.scope}/conv1', f'{u.scope}/conv2'
    f2 = self._bnb_args(c2) is not None   # conv3's dgrad reduces conv2's BN backward, conv2's dgrad conv1's
    f1 = self._bnb_args(c1) is not None
    dr, dsc = self._layer_bwd(f'{u.scope}/conv3', dout, da_masked=dout_masked, bnb_below=c2 if f2 else None)
    dr, _ = self._layer_bwd(c2, dr, da_reduced=f2, bnb_below=c1 if f1 else None)
    if u.has_shortcut_conv:
      dx, _ = self._layer_bwd(c1, dr, da_reduced=f1)
      dx, _ = self._layer_bwd(f'{u.scope}/shortcut', dsc, dx_add=dx, dx_mask=in_mask)
    elif u.stride > 1:
      dsub = torch.empty_like(x)
      ops.maxpool_same_bwd(x, dsc, dsub, 1, u.stride)
      dx, _ = self._layer_bwd(c1, dr, dx_add=dsub, dx_mask=in_mask, da_reduced=f1)
    else:
      dx, _ = self._layer_bwd(c1, dr, dx_add=dsc, dx_mask=in_mask, da_reduced=f1)
    return dx

  # ---- whole network -----------------------------------------------------------------------------------
  def forward_train(self, images):
    """-> post-BN low-res logits fp32 [N, h, w, logits_pitch] (tape recorded for backward)."""
    self.tape = {}
    self.ws.stat.zero_()
    x = self._root_fwd(images)
    for u in arch.units():
      x = self._unit_fwd(x, u)
    f = self._layer_fwd(x, 'feature_extractor/extension/decrease_fdims')
    if self.p.fov:
      f = self._layer_fwd(f, arch.FOV_SCOPE)
    if self.p.psp:
      f = self._psp(f, self._layer_fwd)
    N, h, w, d = f.shape
    au = arch.adaptation_units(d)
    s0 = f'{au[0].scope}/conv1'
    o = self.p.w_off[s0]
    arena = self.p.operand if self.dtype == torch.bfloat16 else self.p.master
    w1 = arena[o:o + 3 * d * d].view(3 * d, 1, 1, d)
    a1 = self._layer_fwd(f, s0, w=w1, nch=3 * d, pad=(0, 0), stride=1, dilation=1, out_hw=(h, w), relu=True)
    pitch = self.hier.logits_pitch
    logits = torch.zeros((N, h, w, pitch), dtype=torch.float32, device=self.dev)
    self.tape['features'] = f
    c0 = 0
    for b, (u, (_, lg)) in enumerate(zip(au, arch.BRANCHES)):
      r = self._layer_fwd(a1[..., b * d:(b + 1) * d], f'{u.scope}/conv2')
      r = self._layer_fwd(r, f'{u.scope}/conv3', relu=True, residual=f)
      scope = f'softmax_classifier/{lg}'
      ck = self.p.by_scope[scope].K
      lz = self._layer_fwd(r, scope, y_f32=True, relu=False)
      logits[..., c0:c0 + ck] = lz  # tiny [N,h,w,ck] placement into the pitched logits buffer
      c0 += ck
    if self.p.upsampling == 'hybrid':
      self.tape['pre_upsampler_logits'] = logits
      logits = self._hybrid_upsampler_fwd(logits)
    return logits

  def _hybrid_upsampler_bwd(self, dlogits):
    """Backward of _hybrid_upsampler_fwd: bias gradient (a per-channel sum, into the layer's dbeta accumulator),
    filter gradient (into the arena, in the stored correlation-kernel layout) and the gradient wrt the
    pre-upsampler logits.  fp32 direct kernels on channel slices of the pitched buffers."""
    x = self.tape['pre_upsampler_logits']
    N, h, w, pitch = x.shape
    dx = torch.zeros_like(x)
    ws = self.ws
    scratch = torch.zeros(64, dtype=torch.float64, device=self.dev)
    c0 = 0
    for sc in arch.UPSAMPLING_SCOPES:
      ck = self.p.by_scope[sc].K
      dy = dlogits[..., c0:c0 + ck].contiguous()
      ops.bn_stats(dy, N * h * w, ck, ck, ws.view(ws.stat, 3, self.p.c_off[sc], ck), scratch[:ck])   # dbias = sum dy
      prm = ops.conv_params((N, h, w, ck), (ck, 3, 3, ck), pad=(1, 1), out_hw=(h, w), x_pitch=pitch, y_pitch=ck,
                            dtype=ops.F32, algo=ops.ALGO_DIRECT, accumulate=True)
      ops.conv2d_wgrad(prm, x[..., c0:c0 + ck], dy, self._wgrad_view(sc, (ck, 3, 3, ck)))
      ops.conv2d_dgrad(prm, dy, self.p.w32(sc), dx[..., c0:c0 + ck])
      self._mark_done(sc, ck * 9 * ck)
      c0 += ck
    return dx

  def backward(self, dlogits):
    """dlogits: fp32 [N, h, w, logits_pitch] gradient wrt the post-BN low-res logits.  Fills the
    gradient arena (conv kernels, gammas, betas)."""
    ws = self.ws
    ws.grads.zero_()  # one memset: every wgrad below accumulates (split-K partial sums) into the arena
    self._side_keep = []
    self._flipped = None
    if self.dtype == torch.bfloat16 and self.conv_algo != ops.ALGO_DIRECT:
      table, arena, views = self.p.flip_plan()
      ops.weights_transpose_flip_batched(self.p.operand, arena, table)  # every dgrad bank, one launch
      self._flipped = views
    self._done = [False] * len(self.p.specs)
    self._tail = len(self.p.specs)
    if self.p.upsampling == 'hybrid':
      dlogits = self._hybrid_upsampler_bwd(dlogits)
    f = self.tape['features']
    N, h, w, d = f.shape
    au = arch.adaptation_units(d)
    da1 = torch.empty((N, h, w, 3 * d), dtype=self.dtype, device=self.dev)
    df = None
    c0 = 0
    for b, (u, (_, lg)) in enumerate(zip(au, arch.BRANCHES)):
      scope = f'softmax_classifier/{lg}'
      ck = self.p.by_scope[scope].K
      dl = dlogits[..., c0:c0 + ck].contiguous()
      c0 += ck
      dr, _ = self._layer_bwd(scope, dl)
      dr, dres = self._layer_bwd(f'{u.scope}/conv3', dr)
      self._layer_bwd(f'{u.scope}/conv2', dr, dx_out=da1[..., b * d:(b + 1) * d])
      df = dres if df is None else ops.add_inplace(df, dres)
    s0 = f'{au[0].scope}/conv1'
    dx, _ = self._layer_bwd(s0, da1, dx_add=df)
    if self.p.psp:
      dx = self._psp_bwd(dx)
    if self.p.fov:
      dx, _ = self._layer_bwd(arch.FOV_SCOPE, dx)
    units = list(arch.units())
    use = self._use_premask()   # (a test flips self.premask between two backward passes over one tape)
    masks = [getattr(self.tape.get(f'{u.scope}/conv3'), 'mask', None) if use else None for u in units]   # of each unit's OUTPUT
    dx, _ = self._layer_bwd('feature_extractor/extension/decrease_fdims', dx, dx_mask=masks[-1])
    for i in range(len(units) - 1, -1, -1):
      dx = self._unit_bwd(dx, units[i], dout_masked=masks[i] is not None, in_mask=masks[i - 1] if i > 0 else None)
    self._root_bwd(dx)
    if self.wgrad_stream is not None:
      torch.cuda.current_stream().wait_stream(self.wgrad_stream)   # join: every filter gradient is in the arena
      self._side_keep = []
    # BN parameter gradients: fp64 accumulators -> fp32 gradient arena
    n = self.p.n_chan_pad
    ws.grads[self.p.n_conv_pad:self.p.n_conv_pad + n] = ws.stat[2 * n:3 * n].to(torch.float32)
    ws.grads[self.p.n_conv_pad + n:self.p.n_conv_pad + 2 * n] = ws.stat[3 * n:4 * n].to(torch.float32)
    return ws.grads

  def _psp_bwd(self, dout):
    """Backward of _psp: the 1x1 fusion conv, then per branch ResizeBilinearGrad -> conv/BN backward ->
    AvgPoolGrad accumulated onto the concatenation's pass-through slice."""
    dcat, _ = self._layer_bwd(arch.PSP_SCOPES[4], dout)       # [N, h, w, 5d]
    N, h, w, d5 = dcat.shape
    d = d5 // 5
    dbottom = dcat[..., :d].contiguous()
    for i, (sc, (kh, kw, pshape)) in enumerate(zip(arch.PSP_SCOPES[:4], self._psp_geom)):
      dc = torch.empty(pshape, dtype=dcat.dtype, device=self.dev)
      ops.resize_bilinear_bwd(dcat[..., (i + 1) * d:(i + 2) * d], dc)
      dpooled, _ = self._layer_bwd(sc, dc)
      ops.avgpool_valid_bwd(dpooled, dbottom, kh, kw, accumulate=True)
    return dbottom

  def loss_and_grad(self, logits, labels, H, W, l2_coef=0.1, grad_scale=1.0):
    """define_losses (TRAIN) + its gradient wrt the low-res logits, one fused kernel + finalize."""
    ws = self.ws
    ws.loss_sums.zero_()
    ws.loss_counts.zero_()
    dlogits = torch.zeros_like(logits)
    # algorithmic HBM bytes: the labels once + low-res logits read + low-res gradient written
    nbytes = sum(v.numel() * v.element_size() for v in labels.values() if v is not None)
    nbytes += 2 * logits.shape[0] * logits.shape[1] * logits.shape[2] * sum(self.hier.head_widths) * 4
    compact = labels.get('bbox_coords') is not None or labels.get('image_vectors') is not None
    if compact and not (self.hier.head_widths == (14, 7, 3) and H >= 2 * logits.shape[1] and W >= 2 * logits.shape[2]):
      # compact weak labels outside the column-walking kernel's reach (Vistas heads): rasterise, then the dense path
      labels = dict(labels)
      if labels.get('bbox_coords') is not None:
        labels['prolabels_per_bbox'] = ops.rasterize_bbox_labels(labels.pop('bbox_coords'), labels.pop('bbox_cids'), H, W)
      if labels.get('image_vectors') is not None:
        labels['prolabels_per_image'] = ops.tile_image_labels(labels.pop('image_vectors'), H, W)
      compact = False
    with self._Timed(self, 'loss_fwd_bwd', nbytes):
      if compact:
        assert labels.get('prolabels_per_bbox') is None and labels.get('prolabels_per_image') is None, \
            'weak labels come either dense or compact, not both'
        ops.loss_fwd_bwd_lists(self.hstruct, logits, H, W, labels.get('prolabels_per_pixel'), labels.get('bbox_coords'),
                               labels.get('bbox_cids'), labels.get('image_vectors'), ws.loss_sums, ws.loss_counts, dlogits)
      else:
        ops.loss_fwd_bwd(self.hstruct, logits, H, W, labels.get('prolabels_per_pixel'),
                         labels.get('prolabels_per_bbox'), labels.get('prolabels_per_image'), ws.loss_sums,
                         ws.loss_counts, dlogits)
    ops.loss_finalize(self.hstruct, ws.loss_sums, ws.loss_counts, l2_coef, grad_scale, dlogits, ws.losses)
    return ws.losses, dlogits
