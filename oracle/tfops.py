"""TensorFlow-1.12 op semantics restated with PyTorch CPU fp32 ops.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED: the ops
below are third-party arithmetic (`tensorflow==1.12.0`, un-vendored); each
function names the reference call site that reaches it and restates the
published TF behaviour.

All tensors at this level are NHWC (as in the reference); convolution kernels
are HWIO (`[kh, kw, Cin, Cout]`), the TF variable layout.
"""

import math

import torch
import torch.nn.functional as F


def same_pad(in_size, k, stride, rate=1):
  """TF 'SAME' padding: out = ceil(in/stride); extra pad goes bottom/right.

  Reached from slim.conv2d / slim.max_pool2d(padding='SAME'),
  code/models/resnet50_extended_model_hierarchical.py:335-353.
  """
  k_eff = k + (k - 1) * (rate - 1)
  out = -(-in_size // stride)
  total = max((out - 1) * stride + k_eff - in_size, 0)
  before = total // 2
  return before, total - before, out


def conv2d(x, w, stride=1, rate=1, padding='SAME'):
  """slim.conv2d without bias/normalizer/activation; x NHWC, w HWIO.

  `padding` is 'SAME', 'VALID' or an explicit (top, bottom, left, right).
  """
  kh, kw = w.shape[0], w.shape[1]
  if padding == 'SAME':
    pt, pb, _ = same_pad(x.shape[1], kh, stride, rate)
    pl, pr, _ = same_pad(x.shape[2], kw, stride, rate)
  elif padding == 'VALID':
    pt = pb = pl = pr = 0
  else:
    pt, pb, pl, pr = padding
  xn = x.permute(0, 3, 1, 2)
  xn = F.pad(xn, (pl, pr, pt, pb))
  wn = w.permute(3, 2, 0, 1)
  y = F.conv2d(xn, wn, stride=stride, dilation=rate)
  return y.permute(0, 2, 3, 1)


def conv2d_same(x, w, stride, rate=1):
  """slim resnet_utils.conv2d_same [TF-1.12]: stride 1 -> SAME; stride > 1 ->
  explicit symmetric-ish zero pad ((k_eff-1)//2 before, rest after) + VALID.

  Reached from resnet_v1.bottleneck conv2 and the root conv1
  (code/models/resnet50_extended_feature_extractor.py:25-30).
  """
  k = w.shape[0]
  if stride == 1:
    return conv2d(x, w, 1, rate, 'SAME')
  k_eff = k + (k - 1) * (rate - 1)
  total = k_eff - 1
  beg = total // 2
  end = total - beg
  return conv2d(x, w, stride, rate, (beg, end, beg, end))


def max_pool_same(x, k, stride):
  """slim.max_pool2d(padding='SAME'): padded cells never win the max.

  code/models/resnet50_extended_model_hierarchical.py:351-353 (arg scope), used
  for resnet pool1 (3x3 s2) and the 1x1 s2 shortcut subsample.
  """
  pt, pb, _ = same_pad(x.shape[1], k, stride)
  pl, pr, _ = same_pad(x.shape[2], k, stride)
  xn = x.permute(0, 3, 1, 2)
  xn = F.pad(xn, (pl, pr, pt, pb), value=float('-inf'))
  y = F.max_pool2d(xn, k, stride)
  return y.permute(0, 2, 3, 1)


def avg_pool_valid(x, ksize, stride):
  """slim.layers.avg_pool2d (default padding='VALID'): windows that do not fit are dropped.
  code/models/resnet50_extended_model_hierarchical.py:191-200 (the PSP bins)."""
  xn = x.permute(0, 3, 1, 2)
  y = F.avg_pool2d(xn, kernel_size=tuple(ksize), stride=tuple(stride), padding=0)
  return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, f_hwoi, bias=None):
  """slim.conv2d_transpose(inputs, C, kernel_size, stride=1, padding='SAME') [TF-1.12]: the gradient of a
  stride-1 SAME conv2d wrt its input; the filter variable has shape [kh, kw, out_channels, in_channels];
  biases are added (the arg scope of the model only configures slim.conv2d, so this layer has no normaliser).
  code/models/resnet50_extended_model_hierarchical.py:168-179 (`--upsampling_method hybrid`)."""
  kh, kw = f_hwoi.shape[0], f_hwoi.shape[1]
  assert kh % 2 == 1 and kw % 2 == 1
  w = f_hwoi.permute(3, 2, 0, 1)   # torch: [in, out, kh, kw]
  y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, stride=1, padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1)
  return y if bias is None else y + bias


def batch_norm(x, gamma, beta, moving_mean, moving_var, training, decay=0.9, eps=1e-5):
  """tf.contrib.layers.batch_norm (fused, NHWC) [TF-1.12].

  training=True: batch mean / biased variance over N*H*W; returns the updated
  moving stats (moving_variance uses the unbiased variance).  training=False:
  affine with the moving stats.
  code/models/resnet50_extended_model_hierarchical.py:298-312,325.
  Returns (y, new_moving_mean, new_moving_var, saved_mean, saved_var).
  """
  if training:
    n = x.shape[0] * x.shape[1] * x.shape[2]
    mean = x.mean(dim=(0, 1, 2))
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    with torch.no_grad():
      unbiased = var * (n / max(n - 1, 1))
      new_mm = moving_mean - (1.0 - decay) * (moving_mean - mean)
      new_mv = moving_var - (1.0 - decay) * (moving_var - unbiased)
    return y, new_mm, new_mv, mean, var
  y = (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta
  return y, moving_mean, moving_var, moving_mean, moving_var


def group_norm(x, gamma, beta, groups, eps=1e-5):
  """tf.contrib.layers.group_norm(x, groups, channels_axis=-1, reduction_axes=(-3, -2), epsilon) [TF-1.12]:
  x [N, H, W, C] -> [N, H, W, G, C/G]; nn.moments over (H, W, C/G) (biased variance); gain = rsqrt(var + eps)
  * gamma, offset = beta - mean * gain; y = x * gain + offset.  No moving statistics.
  code/models/resnet50_extended_model_hierarchical.py:314-333 (`--norm_layer group`)."""
  N, H, W, C = x.shape
  xg = x.reshape(N, H, W, groups, C // groups)
  mean = xg.mean(dim=(1, 2, 4), keepdim=True)
  var = ((xg - mean) ** 2).mean(dim=(1, 2, 4), keepdim=True)
  y = ((xg - mean) * torch.rsqrt(var + eps)).reshape(N, H, W, C)
  return y * gamma + beta


def cross_replica_batch_norm(xs, gamma, beta, moving_mean, moving_var, decay=0.9, eps=1e-5):
  """--cross_replica_norm, training mode: `xs` is the list of per-replica NHWC inputs.
  code/utils/cross_replica_batch_normalization.py:398-459:
    global_mean = sum_r mean_r / R ; global_sq_mean = sum_r E_r[x^2] / R (:404-427, each replica
    divides its moments by num_towers before the SUM reduction); variance = global_sq_mean -
    global_mean^2 (biased); y_r = batch_normalization(x_r, global mean / variance) (:430-431);
    the moving variance takes variance * (n - 1) / n with n the PER-REPLICA sample size - the
    "Bessel removal" that the layer applies although this variance never had the correction
    (:452-459); moving <- moving - (moving - value) * (1 - momentum) (:381-389).
  Gradients flow through the reduced moments (the SUM all-reduce is its own transpose).
  Returns (list of y_r, new_moving_mean, new_moving_var, global_mean, global_var)."""
  R = len(xs)
  mean = sum(x.mean(dim=(0, 1, 2)) / R for x in xs)
  sq = sum((x * x).mean(dim=(0, 1, 2)) / R for x in xs)
  var = sq - mean * mean
  ys = [(x - mean) * torch.rsqrt(var + eps) * gamma + beta for x in xs]
  with torch.no_grad():
    n = xs[0].shape[0] * xs[0].shape[1] * xs[0].shape[2]
    new_mm = moving_mean - (1.0 - decay) * (moving_mean - mean)
    new_mv = moving_var - (1.0 - decay) * (moving_var - var * ((n - 1.0) / n))
  return ys, new_mm, new_mv, mean, var


def _interp_coords(in_size, out_size, align_corners):
  if align_corners and out_size > 1:
    scale = (in_size - 1) / (out_size - 1)
  else:
    scale = in_size / out_size
  return scale


def resize_bilinear(x, out_h, out_w, align_corners=True):
  """tf.image.resize_images(bilinear, align_corners) [TF-1.12 kernel].

  src = dst * scale (no half-pixel offset); lo = floor(src); hi = min(lo+1, in-1);
  lerp in x then y, all in fp32.  The scale is computed in fp32 as TF does
  (`CalculateResizeScale` returns float).
  code/models/resnet50_extended_model_hierarchical.py:167.
  """
  n, h, w, c = x.shape
  sh = torch.tensor(_interp_coords(h, out_h, align_corners), dtype=torch.float32)
  sw = torch.tensor(_interp_coords(w, out_w, align_corners), dtype=torch.float32)
  ys = torch.arange(out_h, dtype=torch.float32) * sh
  xs = torch.arange(out_w, dtype=torch.float32) * sw
  y0 = ys.floor().long()
  x0 = xs.floor().long()
  y1 = torch.clamp(y0 + 1, max=h - 1)
  x1 = torch.clamp(x0 + 1, max=w - 1)
  ty = (ys - y0.float()).view(1, out_h, 1, 1)
  tx = (xs - x0.float()).view(1, 1, out_w, 1)
  top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * tx
  bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * tx
  return top + (bot - top) * ty


def resize_nearest(x, out_h, out_w, align_corners=True):
  """tf.image.resize_images(NEAREST_NEIGHBOR, align_corners=True) [TF-1.12]:
  src = min(roundf(dst*scale), in-1).
  code/estimator/define_estimator_hierarchical.py:559-563.  x is N,H,W[,C].
  """
  h, w = x.shape[1], x.shape[2]
  sh = torch.tensor(_interp_coords(h, out_h, align_corners), dtype=torch.float32)
  sw = torch.tensor(_interp_coords(w, out_w, align_corners), dtype=torch.float32)
  ys = torch.arange(out_h, dtype=torch.float32) * sh
  xs = torch.arange(out_w, dtype=torch.float32) * sw
  if align_corners:
    # roundf: half away from zero (coords are non-negative); floor(x) + (x - floor(x) >= 0.5) is exact in
    # fp32 where floor(x + 0.5) can round a fraction just below one half upwards
    yi = (torch.floor(ys) + (ys - torch.floor(ys) >= 0.5).float()).long()
    xi = (torch.floor(xs) + (xs - torch.floor(xs) >= 0.5).float()).long()
  else:
    yi = torch.floor(ys).long()
    xi = torch.floor(xs).long()
  yi = torch.clamp(yi, max=h - 1)
  xi = torch.clamp(xi, max=w - 1)
  return x[:, yi][:, :, xi]


def softmax(x):
  """tf.nn.softmax over the last axis, fp32."""
  return torch.softmax(x, dim=-1)


def argmax_first(x):
  """tf.argmax(x, -1) cast to int32: lowest index among ties."""
  m = x.max(dim=-1, keepdim=True).values
  c = x.shape[-1]
  idx = torch.arange(c).expand_as(x)
  cand = torch.where(x == m, idx, torch.full_like(idx, c))
  return cand.min(dim=-1).values.to(torch.int32)


def sparse_softmax_cross_entropy(logits, labels):
  """tf.nn.sparse_softmax_cross_entropy_with_logits: -log_softmax(logits)[label]."""
  lsm = torch.log_softmax(logits, dim=-1)
  return -lsm.gather(-1, labels.long().unsqueeze(-1)).squeeze(-1)


def softmax_cross_entropy(logits, soft_labels):
  """tf.nn.softmax_cross_entropy_with_logits (v1): -sum_k t_k log_softmax_k;
  no gradient flows to the labels."""
  lsm = torch.log_softmax(logits, dim=-1)
  return -(soft_labels.detach() * lsm).sum(dim=-1)


def compute_weighted_loss(losses, weights):
  """tf.losses.compute_weighted_loss, default SUM_BY_NONZERO_WEIGHTS with
  safe-div: sum(loss*w) / count(w != 0), 0 if the count is 0.
  code/estimator/define_losses_hierarchical.py:191-199."""
  total = (losses * weights).sum()
  present = (weights != 0).sum().to(torch.float32)
  if present.item() == 0:
    return total * 0.0, present
  return total / present, present


def unsorted_segment_sum_last(data, segment_ids, num_segments):
  """reference `_segment_sum` (define_losses_hierarchical.py:219-224):
  out[..., k] = sum_{c: ids[c]==k} data[..., c]."""
  out = torch.zeros(*data.shape[:-1], num_segments, dtype=data.dtype)
  ids = torch.as_tensor(segment_ids, dtype=torch.long)
  out.index_add_(-1, ids, data)
  return out


def variance_scaling_trunc_normal(shape_hwio, generator):
  """slim.variance_scaling_initializer() contrib defaults [TF-1.12]:
  factor=2, FAN_IN, truncated normal with stddev sqrt(1.3*2/fan_in)
  (values beyond 2 sigma are re-drawn).
  code/models/resnet50_extended_model_hierarchical.py:337."""
  kh, kw, cin, _ = shape_hwio
  std = math.sqrt(1.3 * 2.0 / (kh * kw * cin))
  w = torch.empty(shape_hwio, dtype=torch.float32)
  torch.nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=generator)
  return w
