"""Oracle for one optimizer step of the TRAIN branch of `define_estimator`.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned against the reference-run vectors of
tests/golden/reference_train_run.npz (define_estimator executed in TRAIN mode over tests/golden/tf_shim with the
reference's own model(); tests/test_reference_fixtures.py).

Follows code/estimator/define_estimator_hierarchical.py:77-129:
  :72      model_fn in training mode (batch statistics; the moving-statistic updates go to UPDATE_OPS)
  :84-85   define_losses -> total = segmentation + sum of the slim l2 regularisers (every convolution kernel)
  :96-111  ExponentialMovingAverage(ema_decay, num_updates=global_step).apply over the model variables without
           'BatchNorm/moving' in their names, queued in UPDATE_OPS
  :117     define_optimizer(global_step): the learning rate of the PRE-increment step
  :120-129 create_train_op: UPDATE_OPS (moving statistics, EMA of the variables as they are BEFORE this step's
           update) -> Momentum update of every trainable variable -> global_step += 1
"""

import torch

from oracle import losses as olosses
from oracle import network as onet
from oracle import optimizer as oopt


class TrainState:
  """What the TF session holds between steps: variables, Momentum slots, EMA shadows, the global step."""

  def __init__(self, tf_params, ema_decay=0.0, optimizer='SGDM'):
    self.vars = {k: v.detach().clone() for k, v in tf_params.items()}
    self.trainable = [k for k in self.vars if '/moving_' not in k]
    # tf.train.GradientDescentOptimizer ('SGD', define_optimizer.py:21-22) creates no slot variables
    self.momentum = {k: torch.zeros_like(self.vars[k]) for k in self.trainable} if optimizer == 'SGDM' else {}
    self.ema_decay = float(ema_decay)
    # [TF-1.12] the shadow of a tf.Variable starts at the variable's initial value (no zero-debias for Variables)
    self.ema = {k: self.vars[k].clone() for k in self.trainable} if self.ema_decay > 0 else {}
    self.global_step = 0


def train_step(state, images, labels, dataset, lr, momentum=0.9, nesterov=False, regularization_weight=0.00017,
               bn_decay=0.9, storage='fp32', **model_flags):
  """One `session.run(train_op)`; -> dict of the step's losses (python floats) and the gradients.
  model_flags: psp / fov / upsampling / norm of oracle.network.Net."""
  trainable = set(state.trainable)
  params = {k: v.clone().requires_grad_(k in trainable) for k, v in state.vars.items()}
  net = onet.Net(params, dataset, training=True, bn_decay=bn_decay, storage=storage, **model_flags)
  pred = net.forward(images)
  kernels = [params[k] for k in state.trainable if k.endswith('/weights')]
  losses = olosses.define_losses(pred, labels, dataset, conv_weights=kernels, regularization_weight=regularization_weight)
  losses['total'].backward()
  with torch.no_grad():
    # UPDATE_OPS first: moving statistics of this forward pass, then the EMA of the not-yet-updated variables
    for k, v in net.new_moving.items():
      state.vars[k] = v.detach().clone()
    if state.ema_decay > 0:
      d = oopt.ema_decay_at(state.ema_decay, state.global_step)
      for k in state.trainable:
        state.ema[k] = oopt.ema_update(state.ema[k], state.vars[k], d)
    grads = {}
    for k in state.trainable:
      g = params[k].grad
      grads[k] = g
      if state.momentum:
        state.vars[k], state.momentum[k] = oopt.momentum_step(state.vars[k], g, state.momentum[k], lr, momentum, nesterov)
      else:
        state.vars[k] = state.vars[k] - lr * g
    state.global_step += 1
  out = {k: float(v.detach()) for k, v in losses.items() if k != 'counts'}
  out['grads'] = grads
  return out
