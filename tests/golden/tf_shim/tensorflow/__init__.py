"""Minimal eager emulation of the TensorFlow-1.12 API surface the REFERENCE's own Python touches on the hot path.

Test infrastructure (tests/golden/make_reference_fixtures.py only; never imported by the product, the tests or the
bench): `tensorflow==1.12.0` (code/requirements.txt:9) cannot be installed here (Python 3.12, no network), so the
reference's modules cannot even be imported.  With this package first on sys.path they import and RUN: every
`tf.*` call they make lands in a few lines of torch / numpy below that restate the published TF-1.12 semantics of
that one primitive.  What this buys: the reference's OWN graph-construction code - label
tables, gathers, segment sums, masks, weights, loss normalisation, id remapping, resize calls - is executed, not
re-read, and its outputs (and, through torch autograd, its gradients) become the golden vectors the oracle and the
CUDA path are tested against.  What it does not buy: TensorFlow's kernels themselves are still restated.

Tensors are plain torch tensors (slicing with `...`, arithmetic, comparisons and `.shape[1:]` unpacking behave as
the reference's code expects).  Anything not implemented raises on CALL, never silently returns.
"""

import contextlib
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np
import torch

__version__ = '1.12.0-shim'

float32, float64, int32, int64, uint8, bool_ = torch.float32, torch.float64, torch.int32, torch.int64, torch.uint8, torch.bool
newaxis = None


# ------------------------------------------------------------------------------------------------ stub submodules
class _Missing:
  """Attribute of a stubbed tensorflow submodule: importable, but raises if the reference actually calls it."""

  def __init__(self, name):
    self._name = name

  def __call__(self, *a, **k):
    raise NotImplementedError(f'tf shim: {self._name} is not emulated')

  def __getattr__(self, item):
    if item.startswith('__'):
      raise AttributeError(item)
    return _Missing(f'{self._name}.{item}')


class _StubModule(types.ModuleType):
  def __getattr__(self, item):
    if item.startswith('__'):
      raise AttributeError(item)
    return _Missing(f'{self.__name__}.{item}')


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
  """Every `tensorflow.<anything>` import the reference does at module top level resolves to a stub module."""

  def find_spec(self, fullname, path, target=None):
    if fullname.startswith('tensorflow.'):
      return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
    return None

  def create_module(self, spec):
    known = _SUBMODULES.get(spec.name)
    if known is not None:
      return known
    m = _StubModule(spec.name)
    m.__path__ = []
    return m

  def exec_module(self, module):
    pass


def _sub(name):
  m = _StubModule(name)
  m.__path__ = []
  return m


_SUBMODULES = {}


def _register(name, **attrs):
  m = _sub(name)
  for k, v in attrs.items():
    setattr(m, k, v)
  _SUBMODULES[name] = m
  return m


# ------------------------------------------------------------------------------------------------ tensors with a TF shape
class _Dim(int):
  """tf.Dimension: an int with `.value`."""

  @property
  def value(self):
    return int(self)


class _Shape(tuple):
  """tf.TensorShape of an eager tensor: always fully defined."""

  def __new__(cls, dims):
    return super().__new__(cls, (_Dim(d) for d in dims))

  def __getitem__(self, i):
    r = tuple.__getitem__(self, i)
    return _Shape(r) if isinstance(i, slice) else r

  @property
  def ndims(self):
    return len(self)

  def as_list(self):
    return [int(d) for d in self]

  def is_fully_defined(self):
    return True

  def assert_is_fully_defined(self):
    return None

  def with_rank(self, rank):
    assert len(self) == rank
    return self

  def with_rank_at_least(self, rank):
    assert len(self) >= rank
    return self

  def assert_has_rank(self, rank):
    assert len(self) == rank

  def is_compatible_with(self, other):
    return tuple(self) == tuple(other)


class TFTensor(torch.Tensor):
  """torch.Tensor whose `.shape` behaves like a tf.TensorShape (`.ndims`, `.as_list()`, `shape[-1].value`, ...) and
  that accepts `set_shape` / `get_shape`.  Results of torch operations on it are TFTensors again."""

  @property
  def shape(self):
    return _Shape(torch.Tensor.size(self))

  def get_shape(self):
    return self.shape

  def set_shape(self, shape_):
    for have, want in zip(torch.Tensor.size(self), shape_):
      assert want is None or int(want) == have, f'set_shape({tuple(shape_)}) contradicts {tuple(torch.Tensor.size(self))}'


def as_tf(x):
  """(shim helper) wrap a torch tensor for reference code that inspects TensorShape attributes."""
  return x.as_subclass(TFTensor)


# ------------------------------------------------------------------------------------------------ helpers
def _t(x, dtype=None):
  if isinstance(x, torch.Tensor):
    return x if dtype is None else x.to(dtype)
  return torch.as_tensor(np.asarray(x), dtype=dtype) if dtype is not None else torch.as_tensor(np.asarray(x))


def _int(x):
  return int(x.item()) if isinstance(x, torch.Tensor) else int(x)


# ------------------------------------------------------------------------------------------------ array ops
def cast(x, dtype, name=None):
  return _t(x).to(dtype)


def to_int32(x):
  return _t(x).to(torch.int32)


def constant(value, dtype=None, shape=None, name=None):
  if shape is not None:
    return torch.full(tuple(int(s) for s in shape), value, dtype=dtype)
  return torch.tensor(value, dtype=dtype if dtype is not None else (torch.float32 if isinstance(value, float) else None))


def gather(params, indices, name=None):
  return _t(params)[_t(indices).long()]


def concat(values, axis, name=None):
  return torch.cat([_t(v) for v in values], dim=axis)


def stop_gradient(x, name=None):
  return x.detach()


def one_hot(indices, depth, dtype=torch.float32):
  return torch.nn.functional.one_hot(_t(indices).long(), _int(depth)).to(dtype)


def shape(x, name=None):
  return tuple(x.shape)


def zeros(shape_, dtype=torch.float32):
  return torch.zeros(tuple(int(s) for s in shape_), dtype=dtype)


def zeros_like(x, dtype=None):
  return torch.zeros_like(x, dtype=dtype)


def ones_like(x, dtype=None):
  return torch.ones_like(x, dtype=dtype)


def transpose(x, perm):
  return x.permute(*perm)


def reshape(x, shape_):
  return x.reshape(tuple(shape_))


def identity(x, name=None):
  return x


def where(cond, x, y):
  return torch.where(cond, x, y)


def equal(a, b):
  return _t(a) == (b if not isinstance(b, torch.Tensor) else b)


def greater(a, b):
  return a > b


def greater_equal(a, b):
  return a >= b


def logical_and(a, b):
  return torch.logical_and(a, b)


def reduce_max(x, axis=None):
  x = _t(x)
  return x.max() if axis is None else x.max(dim=axis).values


def reduce_sum(x, axis=None):
  x = _t(x)
  return x.sum() if axis is None else x.sum(dim=axis)


def reduce_mean(x, axis=None):
  x = _t(x)
  return x.mean() if axis is None else x.mean(dim=axis)


def add_n(xs):
  out = xs[0]
  for x in xs[1:]:
    out = out + x
  return out


def div(a, b):
  return a / b


def unstack(x, num=None, axis=0):
  return [int(v) for v in x] if isinstance(x, (tuple, list)) else list(torch.unbind(x, dim=axis))


def stack(values, axis=0):
  return torch.stack([_t(v) for v in values], dim=axis)


def subtract(a, b):
  return _t(a) - _t(b)


def add(a, b):
  return _t(a) + _t(b)


def maximum(a, b):
  return torch.maximum(_t(a, torch.float64) if not isinstance(a, torch.Tensor) else a, _t(b, torch.float64) if not isinstance(b, torch.Tensor) else b)


def minimum(a, b):
  return torch.minimum(_t(a, torch.float64) if not isinstance(a, torch.Tensor) else a, _t(b, torch.float64) if not isinstance(b, torch.Tensor) else b)


def ceil(x):
  return torch.ceil(x)


def assert_equal(a, b, message=None):
  ok = tuple(int(v) for v in a) == tuple(int(v) for v in b)
  assert ok, message or f'tf.assert_equal: {a} != {b}'
  return torch.tensor(True)


RANDOM_LOG = []                      # (shim helper) every value tf.random_uniform returned, in call order
_rng = torch.Generator().manual_seed(0)


def set_random_seed(seed):
  _rng.manual_seed(int(seed))
  del RANDOM_LOG[:]


def random_uniform(shape_, minval=0, maxval=None, dtype=torch.float32, seed=None, name=None):
  """integer dtype: uniform over [minval, maxval)."""
  assert dtype in (torch.int32, torch.int64) and tuple(shape_) == ()
  v = torch.randint(_int(minval), _int(maxval), (), generator=_rng, dtype=dtype)
  RANDOM_LOG.append(int(v))
  return v


def argmax(x, axis):
  # [TF-1.12] lowest index among ties, int64.  torch.argmax does not guarantee first-index on every backend:
  # take the first position equal to the maximum explicitly
  m = x.max(dim=axis, keepdim=True).values
  idx = torch.arange(x.shape[axis]).reshape([-1 if i == (axis % x.dim()) else 1 for i in range(x.dim())])
  big = torch.where(x == m, idx, torch.full_like(idx, x.shape[axis]))
  return big.min(dim=axis).values.to(torch.int64)


def unsorted_segment_sum(data, segment_ids, num_segments):
  """out[k] = sum over rows c of `data` with segment_ids[c] == k (first dimension)."""
  out = torch.zeros((_int(num_segments),) + tuple(data.shape[1:]), dtype=data.dtype)
  return out.index_add(0, _t(segment_ids).long(), data)


def confusion_matrix(labels, predictions, num_classes):
  """[TF-1.12] tf.confusion_matrix: cm[label, prediction] += 1, int32."""
  n = _int(num_classes)
  flat = _t(labels).long() * n + _t(predictions).long()
  return torch.bincount(flat, minlength=n * n).reshape(n, n).to(torch.int32)


def diag_part(x):
  return torch.diagonal(x)


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
  yield str(name)


@contextlib.contextmanager
def control_dependencies(deps):
  for d in deps:
    if isinstance(d, torch.Tensor) and d.dtype == torch.bool:
      assert bool(d.all()), 'tf shim: control dependency (assertion tensor) is false'
  yield


@contextlib.contextmanager
def device(name):
  yield


@contextlib.contextmanager
def variable_scope(name, *a, **k):
  yield str(name)


# ------------------------------------------------------------------------------------------------ tf.nn
def _sparse_ce(labels=None, logits=None, name=None):
  lp = torch.log_softmax(logits, dim=-1)
  return -lp.gather(-1, _t(labels).long().unsqueeze(-1)).squeeze(-1)


def _dense_ce(labels=None, logits=None, name=None, dim=-1):
  # v1: no gradient flows into `labels`
  return -(labels.detach() * torch.log_softmax(logits, dim=dim)).sum(dim=dim)


def _top_k(x, k=1, sorted=True):  # noqa: A002 (TF's keyword)
  # [TF-1.12] top_k: descending, ties broken by LOWER index first
  order = torch.argsort(-x, dim=-1, stable=True)[..., :k]
  return x.gather(-1, order), order.to(torch.int32)


nn = _register('tensorflow.nn', sparse_softmax_cross_entropy_with_logits=_sparse_ce,
               softmax_cross_entropy_with_logits=_dense_ce, softmax=lambda x, name=None: torch.softmax(x, dim=-1),
               top_k=_top_k)


# ------------------------------------------------------------------------------------------------ tf.losses
class _Collections:
  losses = []
  regularization = []


def reset_collections():
  _Collections.losses, _Collections.regularization = [], []


def add_regularization_loss(x):
  """(shim helper) what slim's weights_regularizer would have put into REGULARIZATION_LOSSES."""
  _Collections.regularization.append(x)


def _compute_weighted_loss(losses_, weights=1.0, scope=None, loss_collection='losses', reduction='weighted_sum_by_nonzero_weights'):
  """[TF-1.12] Reduction.SUM_BY_NONZERO_WEIGHTS: sum(losses * weights) / count(broadcast(weights) != 0), 0 if none."""
  assert reduction == 'weighted_sum_by_nonzero_weights'
  w = _t(weights).to(losses_.dtype)
  total = (losses_ * w).sum()
  present = (torch.broadcast_to(w, losses_.shape) != 0).to(losses_.dtype).sum()
  out = torch.where(present > 0, total / torch.where(present > 0, present, torch.ones_like(present)), torch.zeros_like(total))
  if loss_collection is not None:
    _Collections.losses.append(out)
  return out


def _get_total_loss(add_regularization_losses=True, name='total_loss'):
  parts = list(_Collections.losses)
  if add_regularization_losses:
    parts += list(_Collections.regularization)
  return add_n(parts)


losses = _register('tensorflow.losses', compute_weighted_loss=_compute_weighted_loss,
                   add_loss=lambda x, loss_collection=None: _Collections.losses.append(x),
                   get_regularization_losses=lambda scope=None: list(_Collections.regularization),
                   get_total_loss=_get_total_loss)


# ------------------------------------------------------------------------------------------------ tf.image
class _ResizeMethod:
  BILINEAR, NEAREST_NEIGHBOR, BICUBIC, AREA = 0, 1, 2, 3


def _resize_images(images, size, method=_ResizeMethod.BILINEAR, align_corners=False):
  """[TF-1.12] resize_images on NHWC (no half-pixel offset).  Bilinear: fp32, interpolate x first then y (the
  ResizeBilinear kernel's order: top = tl + (tr - tl) * xl, bottom likewise, out = top + (bottom - top) * yl);
  nearest: src = min(roundf(dst * scale), in - 1) when align_corners else floor."""
  x = images
  oh, ow = _int(size[0]), _int(size[1])
  ih, iw = x.shape[1], x.shape[2]

  def scale(i, o):
    return np.float32((i - 1) / (o - 1)) if (align_corners and o > 1) else np.float32(i / o)
  sy, sx = scale(ih, oh), scale(iw, ow)
  if method == _ResizeMethod.NEAREST_NEIGHBOR:
    def idx(o, s, i):
      src = np.arange(o, dtype=np.float32) * s
      r = np.floor(src + np.float32(0.5)) if align_corners else np.floor(src)   # roundf for non-negative values
      return torch.as_tensor(np.minimum(r.astype(np.int64), i - 1))
    return x[:, idx(oh, sy, ih)][:, :, idx(ow, sx, iw)]
  assert method == _ResizeMethod.BILINEAR

  def lerp(o, s, i):
    src = np.arange(o, dtype=np.float32) * s
    lo = np.floor(src).astype(np.int64)
    hi = np.minimum(lo + 1, i - 1)
    return torch.as_tensor(lo), torch.as_tensor(hi), torch.as_tensor((src - lo.astype(np.float32)).astype(np.float32))
  y0, y1, yl = lerp(oh, sy, ih)
  x0, x1, xl = lerp(ow, sx, iw)
  xf = x.to(torch.float32)
  xl = xl.reshape(1, 1, -1, 1)
  yl = yl.reshape(1, -1, 1, 1)
  top = xf[:, y0][:, :, x0] + (xf[:, y0][:, :, x1] - xf[:, y0][:, :, x0]) * xl
  bot = xf[:, y1][:, :, x0] + (xf[:, y1][:, :, x1] - xf[:, y1][:, :, x0]) * xl
  return top + (bot - top) * yl


image = _register('tensorflow.image', resize_images=_resize_images, ResizeMethod=_ResizeMethod)


# ------------------------------------------------------------------------------------------------ tf.train
def _piecewise_constant(x, boundaries, values, name=None):
  """values[0] if x <= b[0]; values[i] if b[i-1] < x <= b[i]; values[-1] beyond."""
  x = _int(x)
  for b, v in zip(boundaries, values):
    if x <= b:
      return float(v)
  return float(values[-1])


def _polynomial_decay(learning_rate, global_step, decay_steps, end_learning_rate=0.0001, power=1.0, cycle=False, name=None):
  step = min(_int(global_step), decay_steps)
  return (learning_rate - end_learning_rate) * (1.0 - step / decay_steps) ** power + end_learning_rate


class _MomentumOptimizer:
  """[TF-1.12] ApplyMomentum: accum = momentum * accum + grad; var -= lr * accum
  (use_nesterov: var -= lr * (grad + momentum * accum))."""

  def __init__(self, learning_rate, momentum, use_locking=False, name='Momentum', use_nesterov=False):
    self.learning_rate, self.momentum, self.use_nesterov = learning_rate, momentum, use_nesterov
    self._learning_rate = learning_rate     # the attribute define_estimator reads for its summaries
    self.slots = {}

  def apply_dense(self, key, var, grad):
    acc = self.slots.get(key)
    if acc is None:
      acc = torch.zeros_like(var)
    acc = self.momentum * acc + grad
    self.slots[key] = acc
    if self.use_nesterov:
      return var - self.learning_rate * (grad + self.momentum * acc)
    return var - self.learning_rate * acc


class _GradientDescentOptimizer:
  def __init__(self, learning_rate, use_locking=False, name='GradientDescent'):
    self.learning_rate = self._learning_rate = learning_rate

  def apply_dense(self, key, var, grad):
    return var - self.learning_rate * grad


class _SessionRunHook:
  """base class only: the reference derives its trace hook from it at import time"""


train = _register('tensorflow.train', SessionRunHook=_SessionRunHook, piecewise_constant=_piecewise_constant, polynomial_decay=_polynomial_decay,
                  MomentumOptimizer=_MomentumOptimizer, GradientDescentOptimizer=_GradientDescentOptimizer)


# ------------------------------------------------------------------------------------------------ misc namespaces
class _ModeKeys:
  TRAIN, EVAL, PREDICT = 'train', 'eval', 'infer'


estimator = _register('tensorflow.estimator', ModeKeys=_ModeKeys)
logging = _register('tensorflow.logging', info=lambda *a, **k: None, warn=lambda *a, **k: None,
                    warning=lambda *a, **k: None, debug=lambda *a, **k: None, INFO=20, DEBUG=10,
                    set_verbosity=lambda *a, **k: None, get_verbosity=lambda: 20, WARN=30)
summary = _register('tensorflow.summary', image=lambda *a, **k: None, scalar=lambda *a, **k: None,
                    histogram=lambda *a, **k: None)


def _deprecated(date, instructions):
  def deco(fn):
    return fn
  return deco


_register('tensorflow.python.util.deprecation', deprecated=_deprecated)

# ------------------------------------------------------------------------------------------------ tf.contrib.slim & co
# (tensorflow/_slim.py: what the reference's model code calls of slim, tf.contrib.layers and slim.nets.resnet_v1)
from tensorflow import _slim  # noqa: E402

variable_scope = _slim.variable_scope   # the stack-keeping version: slim variables are looked up by scope path
nn.relu = _slim.relu
_layers_ns = types.SimpleNamespace(avg_pool2d=_slim.avg_pool2d, batch_norm=_slim.batch_norm, group_norm=_slim.group_norm)
_resnet_v1 = _register('tensorflow.contrib.slim.nets.resnet_v1', resnet_v1_50=_slim.resnet_v1_50, bottleneck=_slim.bottleneck)
_resnet_utils = _register('tensorflow.contrib.slim.nets.resnet_utils', conv2d_same=_slim.conv2d_same, subsample=_slim.subsample)
_nets = _register('tensorflow.contrib.slim.nets', resnet_v1=_resnet_v1, resnet_utils=_resnet_utils)
_slim_mod = _register('tensorflow.contrib.slim', arg_scope=_slim.arg_scope, add_arg_scope=_slim.add_arg_scope, conv2d=_slim.conv2d,
                      conv2d_transpose=_slim.conv2d_transpose, batch_norm=_slim.batch_norm, max_pool2d=_slim.max_pool2d,
                      avg_pool2d=_slim.avg_pool2d, layers=_layers_ns, l2_regularizer=_slim.l2_regularizer,
                      variance_scaling_initializer=_slim.variance_scaling_initializer, nets=_nets)
_framework = _register('tensorflow.contrib.framework', arg_scope=_slim.arg_scope, add_arg_scope=_slim.add_arg_scope)
_layers = _register('tensorflow.contrib.layers', batch_norm=_slim.batch_norm, group_norm=_slim.group_norm,
                    l2_regularizer=_slim.l2_regularizer, variance_scaling_initializer=_slim.variance_scaling_initializer)
_register('tensorflow.contrib', slim=_slim_mod, framework=_framework, layers=_layers)


# ------------------------------------------------------------------------------------------------ training machinery
# (tensorflow/_train.py: global step, UPDATE_OPS, ExponentialMovingAverage, create_train_op, EstimatorSpec)
from tensorflow import _train  # noqa: E402

GraphKeys = _train.GraphKeys
model_variables = _train.model_variables
trainable_variables = _train.trainable_variables
add_to_collection = _train.add_to_collection
get_collection = _train.get_collection
train.get_or_create_global_step = _train.get_or_create_global_step
train.ExponentialMovingAverage = _train.ExponentialMovingAverage
train.Scaffold = _train.Scaffold
train.Saver = _train.Saver
train.list_variables = _train.list_variables
train.init_from_checkpoint = _train.init_from_checkpoint
TensorShape = _train.TensorShape
_register('tensorflow.contrib.distribute.python.values', DistributedValues=_train.DistributedValues,
          TowerLocalVariable=_train.DistributedValues, MirroredVariable=_train.DistributedValues)
global_variables = _train.global_variables
train.SecondOrStepTimer = _train.SecondOrStepTimer
estimator.EstimatorSpec = _train.EstimatorSpec
_register('tensorflow.python.ops', metrics_impl=_register('tensorflow.python.ops.metrics_impl',
                                                         _streaming_confusion_matrix=_train._streaming_confusion_matrix))
_register('tensorflow.contrib.training', create_train_op=_train.create_train_op)
_SUBMODULES['tensorflow.contrib'].training = _SUBMODULES['tensorflow.contrib.training']


class _BaseLayer:
  """base class only: utils/cross_replica_batch_normalization.py derives from tf.layers.BatchNormalization at import time"""

  def __init__(self, *a, **k):
    raise NotImplementedError('tf shim: tf.layers.BatchNormalization is not emulated')


_register('tensorflow.layers', BatchNormalization=_BaseLayer, Layer=_BaseLayer)

sys.meta_path.insert(0, _Finder())
for _name, _m in list(_SUBMODULES.items()):
  sys.modules[_name] = _m


def __getattr__(item):
  if item.startswith('__'):
    raise AttributeError(item)
  full = f'tensorflow.{item}'
  if full in sys.modules:
    return sys.modules[full]
  if item in ('contrib', 'python', 'keras', 'layers', 'data', 'io', 'strings', 'GraphKeys', 'gfile', 'saved_model',
              'initializers', 'distribute', 'metrics', 'test', 'app', 'flags', 'errors', 'dtypes'):
    import importlib
    return importlib.import_module(full)
  return _Missing(full)
