"""Eager emulation of the tf.contrib.slim / tf.contrib.layers / slim.nets.resnet_v1 calls the REFERENCE's model code makes
(code/models/resnet50_extended_model_hierarchical.py, resnet50_extended_feature_extractor.py), so that the reference's
OWN `model()` - its arg scope (eps, decay, scale, is_training), the `feature_extractor` wiring, the three adaptation
bottlenecks, the logits convolutions WITH their normaliser and without activation, `_create_upsampler`,
`_create_psp_module`, softmax / argmax / the decision composition with its literal class-id tables - is EXECUTED on a
parameter dictionary keyed by TF variable names, and the variable names it asks for are recorded.

Test infrastructure (tests/golden/make_reference_fixtures.py only).  tf.contrib.slim is third-party code that is not
part of the reference (TF 1.12, un-vendored): what follows restates its published behaviour - `slim.conv2d` (SAME
padding split, normaliser instead of biases, activation last), `tf.contrib.layers.batch_norm` (fused: biased batch
variance when is_training, moving statistics otherwise), `resnet_utils.conv2d_same` / `subsample` /
`stack_blocks_dense` (output-stride bookkeeping: once the target stride is reached, strides turn into dilation rates),
`resnet_v1.bottleneck` (stride on conv2, `tf.variable_scope(scope, 'bottleneck_v1')`: an explicit scope REPLACES the
default name) and `resnet_v1_50` (blocks 3 / 4 / 6 / 3, the stride on the LAST unit of a block).  It deliberately shares
no code with oracle/ (plain torch.nn.functional calls on NCHW copies).
"""

import contextlib

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------ variable store
VARS = {}            # (shim helper) {TF variable name: torch tensor}, conv kernels HWIO - filled by the fixture script
REQUESTED = []       # (shim helper) every variable name the reference's graph construction asked for, in order
_scope = []          # variable-scope stack
_unique = {}         # default-name counters per enclosing scope ('Conv', 'Conv_1', ...)
UPDATE_OPS = []      # (shim helper) moving-statistic updates a training-mode batch_norm would have queued
NORM_CALLS = []      # (shim helper) (scope, kind, decay, epsilon, scale, is_training / groups) of every normaliser call


def reset(variables=None):
  VARS.clear()
  if variables:
    VARS.update(variables)
  del REQUESTED[:], UPDATE_OPS[:], NORM_CALLS[:], REGULARIZED[:], _scope[:], _arg_stack[:]
  _unique.clear()


def _prefix():
  return '/'.join(_scope)


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, **kw):
  """[TF-1.12] tf.variable_scope(name_or_scope, default_name): default_name is used - uniquified - ONLY when
  name_or_scope is None."""
  if name_or_scope is None:
    key = (_prefix(), default_name)
    n = _unique.get(key, 0)
    _unique[key] = n + 1
    name = default_name if n == 0 else f'{default_name}_{n}'
  else:
    name = str(name_or_scope)
  _scope.append(name)
  try:
    yield '/'.join(_scope)
  finally:
    _scope.pop()


def get_variable(name, shape):
  full = f'{_prefix()}/{name}'
  REQUESTED.append(full)
  if full not in VARS:
    raise KeyError(f'tf shim: the reference asked for variable {full!r} {tuple(shape)}, which the parameter dictionary lacks')
  v = VARS[full]
  assert tuple(v.shape) == tuple(shape), f'{full}: {tuple(v.shape)} != {tuple(shape)}'
  return v


# ------------------------------------------------------------------------------------------------ arg_scope
_arg_stack = []     # stack of {function key: {kwarg: default}}


def _key(fn):
  return getattr(fn, '_shim_key', fn)


@contextlib.contextmanager
def arg_scope(list_ops_or_scope, **kwargs):
  """[TF-1.12] tf.contrib.framework.arg_scope: defaults for the listed ops, nested scopes update the enclosing ones;
  called with a scope object (what `with arg_scope(...) as sc` yielded) it re-enters that scope."""
  if isinstance(list_ops_or_scope, dict):
    assert not kwargs
    new = {k: dict(v) for k, v in list_ops_or_scope.items()}
  else:
    cur = _arg_stack[-1] if _arg_stack else {}
    new = {k: dict(v) for k, v in cur.items()}
    for op in list_ops_or_scope:
      new.setdefault(_key(op), {}).update(kwargs)
  _arg_stack.append(new)
  try:
    yield new
  finally:
    _arg_stack.pop()


def add_arg_scope(fn):
  def wrapped(*args, **kwargs):
    cur = _arg_stack[-1] if _arg_stack else {}
    merged = dict(cur.get(wrapped, {}))
    merged.update(kwargs)
    return fn(*args, **merged)
  wrapped._shim_key = wrapped
  wrapped.__name__ = fn.__name__
  return wrapped


# ------------------------------------------------------------------------------------------------ layers
def _wrap(x):
  from tensorflow import as_tf
  return as_tf(x)


def _same_pad(size, k, stride, rate):
  eff = (k - 1) * rate + 1
  out = -(-size // stride)
  total = max((out - 1) * stride + eff - size, 0)
  return total // 2, total - total // 2


def _pair(v):
  return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list)) or hasattr(v, '__len__') else (int(v), int(v))


def _conv_nhwc(x, w_hwio, stride, rate, padding):
  kh, kw = w_hwio.shape[0], w_hwio.shape[1]
  xn = x.permute(0, 3, 1, 2)
  if padding == 'SAME':
    pt, pb = _same_pad(x.shape[1], kh, stride, rate)
    pl, pr = _same_pad(x.shape[2], kw, stride, rate)
    xn = F.pad(xn, (pl, pr, pt, pb))
  else:
    assert padding == 'VALID'
  y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), stride=stride, dilation=rate)
  return y.permute(0, 2, 3, 1)


@add_arg_scope
def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001, activation_fn=None, is_training=True,
               trainable=True, scope=None, **unused):
  """[TF-1.12] tf.contrib.layers.batch_norm (fused): training = batch mean / BIASED batch variance over N, H, W;
  the moving statistics (updated with the Bessel-corrected variance) only feed inference."""
  with variable_scope(scope, 'BatchNorm'):
    NORM_CALLS.append((_prefix(), 'batch', float(decay), float(epsilon), bool(scale), bool(is_training)))
    c = inputs.shape[-1]
    beta = get_variable('beta', (c,)) if center else torch.zeros(c)
    gamma = get_variable('gamma', (c,)) if scale else torch.ones(c)
    mm, mv = get_variable('moving_mean', (c,)), get_variable('moving_variance', (c,))
    x = torch.Tensor(inputs) if not isinstance(inputs, torch.Tensor) else inputs
    if is_training:
      mean = x.mean(dim=(0, 1, 2))
      var = x.var(dim=(0, 1, 2), unbiased=False)
      n = x.numel() // c
      UPDATE_OPS.append((f'{_prefix()}', mean.detach(), (var * n / max(n - 1, 1)).detach(), decay))
    else:
      mean, var = mm, mv
    y = (x - mean) * torch.rsqrt(var + epsilon) * gamma + beta
    if activation_fn is not None:
      y = activation_fn(y)
  return _wrap(y)


@add_arg_scope
def group_norm(inputs, groups=32, epsilon=1e-6, center=True, scale=True, activation_fn=None, trainable=True, scope=None,
               **unused):
  """[TF-1.12] tf.contrib.layers.group_norm on NHWC: moments per sample over (H, W, channels of a group)."""
  with variable_scope(scope, 'GroupNorm'):
    NORM_CALLS.append((_prefix(), 'group', 0.0, float(epsilon), bool(scale), int(groups)))
    n, h, w, c = inputs.shape
    beta = get_variable('beta', (c,)) if center else torch.zeros(c)
    gamma = get_variable('gamma', (c,)) if scale else torch.ones(c)
    x = inputs.reshape(n, h, w, groups, c // groups)
    mean = x.mean(dim=(1, 2, 4), keepdim=True)
    var = x.var(dim=(1, 2, 4), unbiased=False, keepdim=True)
    y = ((x - mean) * torch.rsqrt(var + epsilon)).reshape(n, h, w, c) * gamma + beta
    if activation_fn is not None:
      y = activation_fn(y)
  return _wrap(y)


def relu(x, name=None):
  return _wrap(torch.relu(x))


@add_arg_scope
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding='SAME', rate=1, activation_fn=relu, normalizer_fn=None,
           normalizer_params=None, weights_initializer=None, weights_regularizer=None, biases_initializer='zeros',
           scope=None, **unused):
  """[TF-1.12] slim.conv2d: convolution, then normalizer_fn(**normalizer_params) INSTEAD of biases, then activation_fn."""
  kh, kw = _pair(kernel_size)
  with variable_scope(scope, 'Conv'):
    cin = int(inputs.shape[-1])
    w = get_variable('weights', (kh, kw, cin, int(num_outputs)))
    if weights_regularizer is not None:
      weights_regularizer(w, f'{_prefix()}/weights')
    y = _conv_nhwc(inputs, w, int(stride), int(rate), padding)
    if normalizer_fn is not None:
      y = normalizer_fn(y, **(normalizer_params or {}))
    elif biases_initializer is not None:
      y = y + get_variable('biases', (int(num_outputs),))
    if activation_fn is not None:
      y = activation_fn(y)
  return _wrap(y)


@add_arg_scope
def conv2d_transpose(inputs, num_outputs, kernel_size, stride=1, padding='SAME', activation_fn=relu, normalizer_fn=None,
                     normalizer_params=None, weights_initializer=None, weights_regularizer=None, biases_initializer='zeros',
                     scope=None, **unused):
  """[TF-1.12] slim.conv2d_transpose, stride 1 / SAME only (what _create_upsampler asks for): filter [kh, kw, out, in],
  y = conv2d_backprop_input, i.e. a correlation with the spatially flipped kernel."""
  kh, kw = _pair(kernel_size)
  assert int(stride) == 1 and padding == 'SAME'
  with variable_scope(scope, 'Conv2d_transpose'):
    cin = int(inputs.shape[-1])
    w = get_variable('weights', (kh, kw, int(num_outputs), cin))
    if weights_regularizer is not None:
      weights_regularizer(w, f'{_prefix()}/weights')
    xn = inputs.permute(0, 3, 1, 2)
    y = F.conv_transpose2d(xn, w.permute(3, 2, 0, 1), padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1)
    if normalizer_fn is not None:
      y = normalizer_fn(y, **(normalizer_params or {}))
    elif biases_initializer is not None:
      y = y + get_variable('biases', (int(num_outputs),))
    if activation_fn is not None:
      y = activation_fn(y)
  return _wrap(y)


@add_arg_scope
def max_pool2d(inputs, kernel_size, stride=2, padding='VALID', scope=None, **unused):
  kh, kw = _pair(kernel_size)
  xn = inputs.permute(0, 3, 1, 2)
  if padding == 'SAME':
    pt, pb = _same_pad(inputs.shape[1], kh, int(stride), 1)
    pl, pr = _same_pad(inputs.shape[2], kw, int(stride), 1)
    xn = F.pad(xn, (pl, pr, pt, pb), value=float('-inf'))
  return _wrap(F.max_pool2d(xn, (kh, kw), stride=int(stride)).permute(0, 2, 3, 1))


@add_arg_scope
def avg_pool2d(inputs, kernel_size, stride=2, padding='VALID', scope=None, **unused):
  kh, kw = _pair(kernel_size)
  sh, sw = _pair(stride)
  assert padding == 'VALID'
  return _wrap(F.avg_pool2d(inputs.permute(0, 3, 1, 2), (kh, kw), stride=(sh, sw)).permute(0, 2, 3, 1))


REGULARIZED = []     # (shim helper) names of the kernels an l2_regularizer was applied to


def l2_regularizer(scale, scope=None):
  """[TF-1.12] slim.l2_regularizer: scale * tf.nn.l2_loss(w) = scale * sum(w^2) / 2, added to REGULARIZATION_LOSSES."""
  def reg(w, name):
    import tensorflow as tf
    REGULARIZED.append((name, float(scale)))
    if w.requires_grad:      # training runs only (the fixture script marks the variables it trains)
      tf.add_regularization_loss(float(scale) * 0.5 * (w * w).sum())
  return reg


def variance_scaling_initializer(*a, **k):
  return 'variance_scaling'


# ------------------------------------------------------------------------------------------------ resnet_utils / resnet_v1
def subsample(inputs, factor, scope=None):
  return inputs if factor == 1 else max_pool2d(inputs, [1, 1], stride=factor, scope=scope)


def conv2d_same(inputs, num_outputs, kernel_size, stride, rate=1, scope=None):
  """[TF-1.12] resnet_utils.conv2d_same: stride 1 -> SAME; else explicit symmetric-ish padding + VALID."""
  if stride == 1:
    return conv2d(inputs, num_outputs, kernel_size, stride=1, rate=rate, padding='SAME', scope=scope)
  eff = kernel_size + (kernel_size - 1) * (rate - 1)
  total = eff - 1
  beg = total // 2
  x = F.pad(inputs.permute(0, 3, 1, 2), (beg, total - beg, beg, total - beg)).permute(0, 2, 3, 1)
  return conv2d(_wrap(x), num_outputs, kernel_size, stride=stride, rate=rate, padding='VALID', scope=scope)


@add_arg_scope
def bottleneck(inputs, depth, depth_bottleneck, stride, rate=1, outputs_collections=None, scope=None,
               use_bounded_activations=False):
  with variable_scope(scope, 'bottleneck_v1'):
    depth_in = int(inputs.shape[-1])
    if depth == depth_in:
      shortcut = subsample(inputs, stride, 'shortcut')
    else:
      shortcut = conv2d(inputs, depth, [1, 1], stride=stride, activation_fn=None, scope='shortcut')
    residual = conv2d(inputs, depth_bottleneck, [1, 1], stride=1, scope='conv1')
    residual = conv2d_same(residual, depth_bottleneck, 3, stride, rate=rate, scope='conv2')
    residual = conv2d(residual, depth, [1, 1], stride=1, activation_fn=None, scope='conv3')
    return relu(shortcut + residual)


def resnet_v1_50(inputs, num_classes=None, is_training=True, global_pool=True, output_stride=None,
                 spatial_squeeze=True, reuse=None, scope='resnet_v1_50'):
  """[TF-1.12] slim.nets.resnet_v1.resnet_v1_50 as a dense feature extractor (num_classes None, no global pool)."""
  assert num_classes is None and not global_pool
  blocks = [('block1', 64, 3, 2), ('block2', 128, 4, 2), ('block3', 256, 6, 2), ('block4', 512, 3, 1)]
  end_points = {}
  bn_scope = arg_scope([batch_norm], is_training=is_training) if is_training is not None else contextlib.nullcontext()
  with variable_scope(scope, 'resnet_v1'), bn_scope:
    assert output_stride is None or output_stride % 4 == 0
    net = conv2d_same(inputs, 64, 7, stride=2, scope='conv1')
    net = max_pool2d(net, [3, 3], stride=2, scope='pool1')
    target = None if output_stride is None else output_stride // 4
    current, rate = 1, 1
    for name, base, units, bstride in blocks:
      with variable_scope(name, 'block'):
        for i in range(units):
          ustride = bstride if i == units - 1 else 1
          with variable_scope('unit_%d' % (i + 1)):
            if target is not None and current == target:
              net = bottleneck(net, base * 4, base, stride=1, rate=rate)
              rate *= ustride
            else:
              net = bottleneck(net, base * 4, base, stride=ustride, rate=1)
              current *= ustride
        end_points[f'{_prefix()}'] = net
    assert target is None or current == target, 'The target output_stride cannot be reached.'
  return net, end_points
