"""Eager emulation of the graph-mode TRAINING machinery the reference's `define_estimator` (TRAIN branch,
code/estimator/define_estimator_hierarchical.py:77-159) assembles: global step, the UPDATE_OPS collection,
tf.train.ExponentialMovingAverage, tf.contrib.training.create_train_op, EstimatorSpec / Scaffold containers.

Test infrastructure (tests/golden/make_reference_train_fixtures.py only).  TensorFlow is un-vendored: what follows restates
its published behaviour -
  * create_train_op(total_loss, optimizer, global_step): every op of GraphKeys.UPDATE_OPS runs BEFORE the gradients are
    applied (with_dependencies on total_loss), gradients are taken w.r.t. tf.trainable_variables(), apply_gradients
    increments the global step last;
  * ExponentialMovingAverage(decay, num_updates, zero_debias).apply(var_list): decay' = min(decay, (1 + n) / (10 + n)),
    shadow <- shadow - (1 - decay') (shadow - var); the shadow of a tf.Variable starts at the variable's initial value and is
    NOT zero-debiased (zero_debias only applies to plain tensors); shadow names <scope>/<var>/ExponentialMovingAverage;
  * fused batch norm's moving-statistic updates (queued by tensorflow/_slim.py::batch_norm) are UPDATE_OPS too.
The reference's own choices - which variables get an EMA, decay and num_updates, the variable scopes, which loss is
differentiated, the optimizer and its schedule - are executed, not restated.
"""

import types

import torch

from tensorflow import _slim

GLOBAL_STEP = [None]
EMA_SHADOWS = {}     # {shadow variable name: tensor}
COLLECTIONS = {}     # user collections (tf.add_to_collection)
METRIC_VARS = {}     # {name: float64 tensor} - local "metric variables" of the evaluation graph (total_confusion_matrix)
SLOT_NAMES = []      # names of the Momentum slot variables create_train_op made: <variable scope>/<variable>/Momentum
INIT_FROM_CHECKPOINT = []   # (checkpoint path, {checkpoint name: graph variable}) of every tf.train.init_from_checkpoint call
CHECKPOINT_VARIABLES = {}   # {checkpoint path: [(name, shape)]} - what tf.train.list_variables reports (set by the script)
OPT_SLOTS = {}       # {variable name: Momentum accumulator} - slot variables outlive the optimizer OBJECT, which the eager
                     # run re-creates on every step (in TF they are graph variables `train_ops/<var>/Momentum`)


class GraphKeys:
  UPDATE_OPS = 'update_ops'
  REGULARIZATION_LOSSES = 'regularization_losses'
  GLOBAL_VARIABLES = 'variables'
  TRAINABLE_VARIABLES = 'trainable_variables'
  MODEL_VARIABLES = 'model_variables'


def reset():
  GLOBAL_STEP[0] = None
  EMA_SHADOWS.clear()
  COLLECTIONS.clear()
  OPT_SLOTS.clear()
  METRIC_VARS.clear()
  del SLOT_NAMES[:], INIT_FROM_CHECKPOINT[:]


class _Var:
  """What tf.model_variables() hands out: `.name` ('<scope>:0'), `.op.name`, and the live tensor."""

  def __init__(self, name):
    self.name = name + ':0'
    self.op = types.SimpleNamespace(name=name)
    self.key = name

  @property
  def value(self):
    return _slim.VARS[self.key]

  @property
  def shape(self):
    base = self.key
    for suffix in ('/Momentum', '/ExponentialMovingAverage'):      # a slot / shadow has its variable's shape
      if base.endswith(suffix):
        base = base[:-len(suffix)].split('/', 1)[1]
    return tuple(_slim.VARS[base].shape) if base in _slim.VARS else ()


_VAR_OBJECTS = {}    # one object per variable name (the savers compare variables by identity)


def _var(name):
  if name not in _VAR_OBJECTS:
    _VAR_OBJECTS[name] = _Var(name)
  return _VAR_OBJECTS[name]


def model_variables():
  seen, out = set(), []
  for n in _slim.REQUESTED:      # slim registers every variable it creates as a model variable, in creation order
    if n not in seen:
      seen.add(n)
      out.append(_var(n))
  return out


def global_variables():
  """Creation order: model variables, the global step and - in the TRAIN graph - the EMA shadows and Momentum slots."""
  return model_variables() + [_var('global_step')] + [_var(n) for n in EMA_SHADOWS] + [_var(n) for n in SLOT_NAMES]


class TensorShape:
  """[TF-1.12] fully defined shapes: compatible = same rank and equal dimensions."""

  def __init__(self, dims):
    self.dims = tuple(int(d) for d in dims)

  def is_compatible_with(self, other):
    return self.dims == tuple(int(d) for d in getattr(other, 'dims', other))


def list_variables(path):
  return list(CHECKPOINT_VARIABLES[path])


def init_from_checkpoint(path, assignment_map):
  INIT_FROM_CHECKPOINT.append((path, dict(assignment_map)))


class DistributedValues:
  """tf.contrib.distribute values container: nothing here is one (single tower)."""


class Saver:
  """Container: the reference's savers only choose WHICH variables are saved / restored under WHICH checkpoint names."""

  def __init__(self, var_list=None, sharded=False, max_to_keep=5, save_relative_paths=False, **kw):
    self.var_list = var_list


def trainable_variables():
  return [v for v in model_variables() if '/moving_' not in v.key]


def get_or_create_global_step():
  if GLOBAL_STEP[0] is None:
    GLOBAL_STEP[0] = torch.zeros((), dtype=torch.int64)
  return GLOBAL_STEP[0]


def add_to_collection(name, value):
  COLLECTIONS.setdefault(name, []).append(value)


def get_collection(name, scope=None):
  return list(COLLECTIONS.get(name, []))


class ExponentialMovingAverage:
  def __init__(self, decay, num_updates=None, zero_debias=False, name='ExponentialMovingAverage'):
    self.decay, self.num_updates, self.name = float(decay), num_updates, name
    self.scope = _slim._prefix()

  def apply(self, var_list=None):
    scope = _slim._prefix()
    names = []
    for v in var_list:
      shadow = f'{scope}/{v.key}/{self.name}'
      # a Variable's shadow starts at its initial value; no zero-debias for Variables
      EMA_SHADOWS.setdefault(shadow, v.value.detach().clone())
      names.append((shadow, v.key))

    def op():
      d = self.decay
      if self.num_updates is not None:
        n = float(int(self.num_updates))
        d = min(d, (1.0 + n) / (10.0 + n))
      for shadow, key in names:
        s = EMA_SHADOWS[shadow]
        s -= (1.0 - d) * (s - _slim.VARS[key].detach())
    return op


def create_train_op(total_loss, optimizer, global_step=None, update_ops=None, variables_to_train=None, summarize_gradients=False,
                    check_numerics=True, **unused):
  assert update_ops is None and variables_to_train is None
  if hasattr(optimizer, 'slots'):
    optimizer.slots = OPT_SLOTS
  variables = trainable_variables()
  # [TF-1.12] slot_creator: apply_gradients creates one slot per trainable variable, named <current variable scope>/
  # <variable op name>/<optimizer name> - the scope is whatever `with tf.variable_scope(...)` the reference wrapped around
  if hasattr(optimizer, 'slots'):
    SLOT_NAMES.extend(f'{_slim._prefix()}/{v.key}/Momentum' for v in variables)
  queued = list(get_collection(GraphKeys.UPDATE_OPS))
  moving = list(_slim.UPDATE_OPS)

  def train_op():
    grads = torch.autograd.grad(total_loss, [v.value for v in variables], allow_unused=True)
    # 1. UPDATE_OPS (moving statistics of this forward pass, then whatever the model function queued: the EMA)
    for scope, mean, var, decay in moving:
      mm, mv = _slim.VARS[f'{scope}/moving_mean'], _slim.VARS[f'{scope}/moving_variance']
      with torch.no_grad():
        mm -= (1.0 - decay) * (mm - mean.detach())
        mv -= (1.0 - decay) * (mv - var.detach())
    with torch.no_grad():
      for op in queued:
        op()
      # 2. apply_gradients, 3. global step
      for v, g in zip(variables, grads):
        if g is None:
          continue
        new = optimizer.apply_dense(v.key, v.value.detach(), g)
        v.value.copy_(new)
      if global_step is not None:
        global_step.add_(1)
    return total_loss.detach()
  train_op.optimizer = optimizer
  return train_op


def EstimatorSpec(mode, predictions=None, loss=None, train_op=None, eval_metric_ops=None, training_hooks=None, scaffold=None,
                  **kw):
  return types.SimpleNamespace(mode=mode, predictions=predictions, loss=loss, train_op=train_op,
                               eval_metric_ops=eval_metric_ops, training_hooks=training_hooks, scaffold=scaffold)


def Scaffold(saver=None, **kw):
  return types.SimpleNamespace(saver=saver)


class SecondOrStepTimer:
  def __init__(self, every_secs=None, every_steps=None):
    self.every_secs, self.every_steps = every_secs, every_steps


def _streaming_confusion_matrix(labels, predictions, num_classes, weights=None):
  """[TF-1.12] tensorflow/python/ops/metrics_impl.py::_streaming_confusion_matrix: a float64 [num_classes, num_classes]
  metric variable `total_confusion_matrix`; labels and predictions are cast to int64 and flattened; update_op =
  assign_add(total_cm, confusion_matrix(labels, predictions, num_classes, dtype=float64)); returns (total_cm, update_op).
  Eager emulation: one define_estimator call stands for one session.run of the evaluation loop, which runs update_op -
  the update is applied here and the returned tensor is the variable AFTER it (what Estimator.evaluate reads at the end)."""
  import tensorflow as tf
  assert weights is None
  num_classes = int(num_classes)
  total = METRIC_VARS.setdefault('total_confusion_matrix', torch.zeros(num_classes, num_classes, dtype=torch.float64))
  lab = torch.as_tensor(labels).to(torch.int64).reshape(-1)
  pred = torch.as_tensor(predictions).to(torch.int64).reshape(-1)
  if int(lab.min()) < 0 or int(lab.max()) >= num_classes or int(pred.min()) < 0 or int(pred.max()) >= num_classes:
    raise ValueError('InvalidArgumentError: confusion_matrix index out of bounds')   # TF fails the run
  total += torch.as_tensor(tf.confusion_matrix(lab, pred, num_classes)).to(torch.float64)
  return total, (lambda: total)
