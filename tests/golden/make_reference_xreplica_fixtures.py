"""Golden vectors of --cross_replica_norm produced by running the reference's own layer code:
tests/golden/reference_xreplica_run.npz.

    python tests/golden/make_reference_xreplica_fixtures.py       # needs /root/reference; run in the build container

`code/utils/cross_replica_batch_normalization.py::CrossReplicaBatchNormalization._fused_batch_norm` (:393-476) and its
`_assign_moving_average` (:381-389) are the reference's OWN functions (the class derives from tf.layers.
BatchNormalization, the rest of which is TensorFlow).  They are called UNMODIFIED, in training mode, once per replica of
a 2-tower MirroredStrategy emulation:
  replica_context.merge_call(_merge_fn, mean / num_towers, square_mean / num_towers) runs the reference's `_merge_fn`,
  whose strategy.reduce(VariableAggregation.SUM, ...) is the SUM over the towers' values (two passes: the first collects
  every tower's arguments - tensors that stay connected to their inputs for autograd -, the second hands out the sums).
Restated TF calls (bound below onto the names the reference module imported): tf.reduce_mean, tf.square,
tf.nn.batch_normalization ((x - mean) * rsqrt(var + eps) * gamma + beta), smart_cond on a python bool, cast, size,
assign_sub.
Stored: per-tower inputs and outputs, the global moments, the moving statistics after the update (with the reference's
(n - 1) / n factor on a variance that never had Bessel's correction, n = the PER-TOWER sample size), and the gradients of
sum_r <y_r, w_r> with respect to every tower's input, gamma and beta (through the cross-tower moments).
tests/test_reference_fixtures.py checks oracle/tfops.py::cross_replica_batch_norm against it; the product's NCCL path is
checked against the same oracle formula in tests/test_gpu_cross_replica.py (2 GPUs).
"""

import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('WLSEG_REFERENCE', '/root/reference/code')
OUT = os.path.join(HERE, 'reference_xreplica_run.npz')
SEED, TOWERS, SHAPE, MOMENTUM, EPSILON = 41, 2, (2, 5, 7, 16), 0.9, 1e-5


class _Variable:
  """A tf.Variable as far as `_assign_moving_average` touches it."""

  def __init__(self, t):
    self.t = t
    self.dtype = types.SimpleNamespace(base_dtype=t.dtype)

  def __sub__(self, other):
    return self.t - other


class _PerTower:
  def __init__(self, slot):
    self.slot = slot


class _Strategy:
  def __init__(self, collected):
    self.collected = collected

  def reduce(self, aggregation, value, destinations=None):
    assert aggregation == 'SUM'
    return sum(args[value.slot] for args in self.collected)


class _TowerContext:
  def __init__(self, num_towers):
    self.num_towers = num_towers
    self.collected = []       # pass 1: every tower's merge_call arguments
    self.collecting = True

  def merge_call(self, merge_fn, *args):
    if self.collecting:
      self.collected.append(args)
      return args
    return merge_fn(_Strategy(self.collected), *[_PerTower(i) for i in range(len(args))])


def main():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import tensorflow as tf
  assert tf.__version__.endswith('shim')
  from utils import cross_replica_batch_normalization as xr

  def cast(x, dtype):
    return torch.as_tensor(x, dtype=getattr(dtype, 'base_dtype', dtype))

  def assign_sub(variable, delta, name=None):
    variable.t = variable.t - delta.detach()
    return variable.t

  xr.tf_utils.smart_cond = lambda pred, true_fn, false_fn: true_fn() if pred else false_fn()
  xr.tf_utils.constant_value = lambda x: x
  xr.math_ops.cast = cast
  xr.array_ops.size = lambda x: x.numel()
  xr.ops.convert_to_tensor = lambda x, name=None: torch.as_tensor(x)
  xr.ops.name_scope = lambda *a, **k: contextlib.nullcontext('AssignMovingAvg')
  xr.state_ops.assign_sub = assign_sub
  tf.square = lambda x: x * x
  tf.nn.batch_normalization = lambda x, mean, variance, offset, scale, eps: (x - mean) * torch.rsqrt(variance + eps) * scale + offset
  tf.VariableAggregation = types.SimpleNamespace(SUM='SUM')
  context = _TowerContext(TOWERS)
  tf.contrib.distribute = types.SimpleNamespace(get_tower_context=lambda: context)

  g = torch.Generator().manual_seed(SEED)
  C = SHAPE[-1]
  xs = [(torch.randn(SHAPE, generator=g) * (1.0 + 0.5 * r) + 0.3 * r).requires_grad_(True) for r in range(TOWERS)]
  ws = [torch.randn(SHAPE, generator=g) for _ in range(TOWERS)]
  gamma = (0.75 + 0.5 * torch.rand(C, generator=g)).requires_grad_(True)
  beta = (0.1 * torch.randn(C, generator=g)).requires_grad_(True)
  moving_mean0 = 0.1 * torch.randn(C, generator=g)
  moving_var0 = 0.75 + 0.5 * torch.rand(C, generator=g)

  def layer():
    updates = []
    return types.SimpleNamespace(
        beta=beta, gamma=gamma, center=True, scale=True, epsilon=EPSILON, momentum=MOMENTUM, _data_format='NHWC',
        _bessels_correction_test_only=False, moving_mean=_Variable(moving_mean0.clone()),
        moving_variance=_Variable(moving_var0.clone()), add_update=lambda u, inputs=None: updates.append(u),
        _assign_moving_average=lambda v, value, m: xr.CrossReplicaBatchNormalization._assign_moving_average(None, v, value, m))

  run = xr.CrossReplicaBatchNormalization._fused_batch_norm
  for x in xs:                                  # pass 1: collect
    run(layer(), tf.as_tf(x), True)
  context.collecting = False
  layers = [layer() for _ in xs]
  ys = [run(l, tf.as_tf(x), True) for l, x in zip(layers, xs)]     # pass 2: the towers' real outputs
  loss = sum((torch.Tensor(y) * w).sum() for y, w in zip(ys, ws))
  grads = torch.autograd.grad(loss, xs + [gamma, beta])
  out = {'momentum': np.asarray(MOMENTUM), 'epsilon': np.asarray(EPSILON), 'gamma': gamma.detach().numpy(),
         'beta': beta.detach().numpy(), 'moving_mean_before': moving_mean0.numpy(), 'moving_variance_before': moving_var0.numpy(),
         'dgamma': grads[-2].numpy(), 'dbeta': grads[-1].numpy()}
  for r in range(TOWERS):
    out[f'tower{r}/x'] = xs[r].detach().numpy()
    out[f'tower{r}/w'] = ws[r].numpy()
    out[f'tower{r}/y'] = torch.Tensor(ys[r]).detach().numpy()
    out[f'tower{r}/dx'] = grads[r].numpy()
    out[f'tower{r}/moving_mean_after'] = layers[r].moving_mean.t.numpy()
    out[f'tower{r}/moving_variance_after'] = layers[r].moving_variance.t.numpy()
  # every tower computes the same update (the moments are global)
  assert np.array_equal(out['tower0/moving_mean_after'], out['tower1/moving_mean_after'])
  assert np.array_equal(out['tower0/moving_variance_after'], out['tower1/moving_variance_after'])
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes; y0 mean', float(out['tower0/y'].mean()))


if __name__ == '__main__':
  main()
