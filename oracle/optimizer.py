"""Oracle for LR schedules, momentum SGD, EMA and the derived step settings.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned against the reference-run vectors
(define_optimizer executed over tests/golden/tf_shim; tests/test_reference_fixtures.py).
"""


def piecewise_constant(step, boundaries, values):
  """[TF-1.12] tf.train.piecewise_constant via code/estimator/define_optimizer.py:5-7:
  values[0] if step <= b[0]; values[i] if b[i-1] < step <= b[i]; values[-1] beyond."""
  assert len(values) == len(boundaries) + 1
  for b, v in zip(boundaries, values):
    if step <= b:
      return v
  return values[-1]


def polynomial_decay(lr0, step, decay_steps, end_lr, power):
  """[TF-1.12] tf.train.polynomial_decay (cycle=False), define_optimizer.py:9-13."""
  s = min(step, decay_steps)
  return (lr0 - end_lr) * (1.0 - s / decay_steps) ** power + end_lr


def train_schedule(Ntrain=2975, Nb=4, Ne=17, boundaries_epochs=(8, 15, 17), lr0=0.01,
                   decay=None, values=None):
  """code/system_factory.py:197-233: steps per epoch, total steps, LR boundaries in
  steps (last boundary popped when Ne - boundaries[-1] == 0) and plateau values."""
  num_batches_per_epoch = int(Ntrain / Nb)
  num_training_steps = int(Ne * num_batches_per_epoch)
  b = list(boundaries_epochs)
  if not (decay or values):
    decay = 0.5
  last = Ne - b[-1]
  if last == 0:
    b.pop()
  elif last < 0:
    raise ValueError('Ne is less than learning rate boundaries.')
  b_steps = [x * num_batches_per_epoch for x in b]
  if decay:
    values = [lr0 * decay ** i for i in range(len(b_steps) + 1)]
  return dict(num_batches_per_epoch=num_batches_per_epoch, num_training_steps=num_training_steps,
              boundaries=b_steps, values=list(values))


def momentum_step(w, g, acc, lr, momentum=0.9, nesterov=False):
  """[TF-1.12] ApplyMomentum via define_optimizer.py:17-20:
  acc <- m*acc + g ; w <- w - lr*acc  (Nesterov: w <- w - lr*(g + m*acc))."""
  acc = momentum * acc + g
  if nesterov:
    w = w - lr * (g + momentum * acc)
  else:
    w = w - lr * acc
  return w, acc


def ema_decay_at(decay, num_updates):
  """[TF-1.12] ExponentialMovingAverage(num_updates): min(decay, (1+t)/(10+t)).
  code/estimator/define_estimator_hierarchical.py:96-111."""
  return min(decay, (1.0 + num_updates) / (10.0 + num_updates))


def ema_update(shadow, var, decay):
  """shadow <- shadow - (1-decay)*(shadow - var)."""
  return shadow - (1.0 - decay) * (shadow - var)
