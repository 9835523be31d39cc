"""Checkpoint interop: variable NAMES, selection and matching rules of the reference's savers / initialisers.

Mirrors
  code/estimator/define_savers.py:3-36     `train_saver`     every global variable is saved
  code/estimator/define_savers.py:38-66    `predict_saver`   checkpoint key -> model variable; with
                                           --restore_emas the key of every model variable that is not a
                                           BatchNorm moving statistic becomes
                                           exponential_moving_averages/<name>/ExponentialMovingAverage
  code/estimator/define_initializers.py:72-131  `replace_initializers`: ImageNet warm start - a checkpoint
                                           variable initialises the model variable whose name CONTAINS its
                                           name (and whose shape is compatible), unless the model
                                           variable's name contains one of the exclude words
  code/estimator/define_estimator_hierarchical.py:96-111   which variables have EMA shadows

The binary container differs: TF-1.12 writes tensor bundles, which need TensorFlow to read or write; here a
checkpoint is a flat {TF variable name: array} dict stored as a torch file (`model.ckpt-<step>.pt`) or a numpy
`.npz` - what `tf.train.load_checkpoint(path)` + `get_tensor(name)` yields for a reference checkpoint (the
three-line export script is in INTEGRATION.md).  Tensors keep TF's layout (conv kernels HWIO).
Host-side bookkeeping only - no arithmetic.
"""

import collections
import os

import numpy as np
import torch

EMA_SCOPE = 'exponential_moving_averages'
EMA_SUFFIX = 'ExponentialMovingAverage'
# tf.train.MomentumOptimizer slot variables are created under the `train_ops` variable scope
# (define_estimator_hierarchical.py:114-117) as <scope>/<variable name>/Momentum  [TF-1.12 slot_creator]
MOMENTUM_SCOPE = 'train_ops'
MOMENTUM_SUFFIX = 'Momentum'
GLOBAL_STEP = 'global_step'

# define_initializers.py:99-104 (`replace_initializers`, the one the TRAIN branch calls)
INIT_EXCLUDE = ('global_step', 'train_ops', 'ExponentialMovingAverage', 'Momentum', 'classifier', 'extension')


def ema_name(var_name):
  return f'{EMA_SCOPE}/{var_name}/{EMA_SUFFIX}'


def momentum_name(var_name):
  return f'{MOMENTUM_SCOPE}/{var_name}/{MOMENTUM_SUFFIX}'


def model_variables(params):
  """[(TF name, TF shape)] of tf.model_variables() in creation order: per convolution `weights` (HWIO) then
  its BatchNorm beta, gamma, moving_mean, moving_variance (slim creates beta before gamma)."""
  out = []
  plain = getattr(params, 'plain', ())
  for s in params.specs:
    if s.scope in plain:   # slim.conv2d_transpose of --upsampling_method hybrid: [kh, kw, out, in] + biases, no BN
      out += [(f'{s.scope}/weights', (s.R, s.S, s.K, s.C)), (f'{s.scope}/biases', (s.K,))]
      continue
    out.append((f'{s.scope}/weights', (s.R, s.S, s.C, s.K)))
    if getattr(params, 'norm', 'batch') == 'group':   # tf.contrib.layers.group_norm: beta, gamma, no moving statistics
      out += [(f'{s.scope}/GroupNorm/beta', (s.K,)), (f'{s.scope}/GroupNorm/gamma', (s.K,))]
      continue
    for v in ('beta', 'gamma', 'moving_mean', 'moving_variance'):
      out.append((f'{s.scope}/BatchNorm/{v}', (s.K,)))
  return out


def has_ema(var_name):
  """define_estimator_hierarchical.py:103-106: every model variable except the BN moving statistics."""
  return 'BatchNorm/moving' not in var_name


def trainable(var_name):
  return has_ema(var_name)


# ---- files ------------------------------------------------------------------------------------------------
def save_file(path, variables, global_step):
  """{name: tensor} + global_step -> .pt (torch) or .npz (numpy) by extension."""
  os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
  if path.endswith('.npz'):
    arrays = {k: np.asarray(v) for k, v in variables.items()}
    arrays[GLOBAL_STEP] = np.asarray(int(global_step), dtype=np.int64)
    np.savez(path, **arrays)
  else:
    torch.save({'global_step': int(global_step), 'variables': {k: torch.as_tensor(v) for k, v in variables.items()}}, path)
  return path


def load_file(path):
  """-> ({name: torch tensor}, global_step).  Accepts the two containers `save_file` writes and a bare
  {name: array} dict saved with torch.save (e.g. an exported ImageNet checkpoint)."""
  if path.endswith('.npz'):
    with np.load(path) as z:
      variables = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
  else:
    blob = torch.load(path, map_location='cpu')
    variables = dict(blob['variables']) if isinstance(blob, dict) and 'variables' in blob else dict(blob)
    if isinstance(blob, dict) and 'global_step' in blob and GLOBAL_STEP not in variables:
      variables[GLOBAL_STEP] = torch.as_tensor(int(blob['global_step']))
    variables = {k: torch.as_tensor(v) for k, v in variables.items()}
  step = int(variables.pop(GLOBAL_STEP)) if GLOBAL_STEP in variables else 0
  return variables, step


# ---- train_saver --------------------------------------------------------------------------------------------
def export_train_state(params, trainer=None):
  """Every global variable of the TRAIN graph under its TF name: model variables, their Momentum slots and,
  with --ema_decay > 0, their ExponentialMovingAverage shadows (define_savers.py:14-32 saves them all)."""
  out = collections.OrderedDict(params.to_tf_dict())
  if trainer is None:
    return out
  # --optimizer SGD: tf.train.GradientDescentOptimizer has no slot variables (define_optimizer.py:21-22)
  arenas = [(momentum_name, trainer.ws.momentum)] if getattr(trainer, 'optimizer', 'SGDM') == 'SGDM' else []
  if getattr(trainer.ws, 'ema_shadow', None) is not None:
    arenas.append((ema_name, trainer.ws.ema_shadow))
  for namer, arena in arenas:
    for name, t in params.arena_to_tf_dict(arena).items():
      out[namer(name)] = t
  return out


def import_train_state(params, trainer, variables):
  """Continue training from a log_dir checkpoint: model variables, and the optimizer / EMA slots when the
  checkpoint has them (a slot that is missing keeps its fresh initial value, as after a warm start)."""
  params.load_tf_dict(variables)
  if trainer is None:
    return
  arenas = [(momentum_name, trainer.ws.momentum)]
  if getattr(trainer.ws, 'ema_shadow', None) is not None:
    trainer.ws.ema_shadow.copy_(params.master)   # TF initialises a shadow with its variable
    arenas.append((ema_name, trainer.ws.ema_shadow))
  for namer, arena in arenas:
    named = {name: variables[namer(name)] for name, _ in model_variables(params)
             if trainable(name) and namer(name) in variables}
    params.load_into_arena(arena, named)


# ---- predict_saver ------------------------------------------------------------------------------------------
def predict_var_dict(params, restore_emas=False):
  """{checkpoint key: model variable name}, define_savers.py:44-56."""
  out = collections.OrderedDict()
  for name, _ in model_variables(params):
    key = ema_name(name) if (restore_emas and has_ema(name)) else name
    out[key] = name
  return out


def select_for_predict(params, variables, restore_emas=False):
  """The {model variable name: tensor} dict an EVAL / PREDICT graph restores from `variables`; a missing key
  raises KeyError naming it (tf.train.Saver.restore fails with NotFoundError there)."""
  out = {}
  for key, name in predict_var_dict(params, restore_emas).items():
    if key not in variables:
      raise KeyError(f'Key {key} not found in checkpoint' + (' (was it trained with --ema_decay > 0?)' if restore_emas else ''))
    out[name] = variables[key]
  return out


# ---- replace_initializers (ImageNet warm start) ---------------------------------------------------------------
def match_init_checkpoint(ckpt_vars, graph_vars, psp_module=False):
  """define_initializers.py:92-115.  ckpt_vars: [(name, shape)] as tf.train.list_variables returns them;
  graph_vars: [(name, shape)] of tf.global_variables().  -> {checkpoint name: graph variable name}.
  A graph variable is skipped when its name contains an exclude word (+ 'psp' without --psp_module); a
  checkpoint variable initialises the graph variable whose name contains its name and whose shape matches
  (the last such graph variable wins, as the reference's dict assignment does)."""
  exclude = list(INIT_EXCLUDE)
  if not psp_module:
    exclude.append('psp')
  var_dict = collections.OrderedDict()
  for gname, gshape in graph_vars:
    gfull = gname + ':0'   # tf.Variable.name carries the output index
    if any(exc in gfull for exc in exclude):
      continue
    for cname, cshape in ckpt_vars:
      if cname in gfull and tuple(cshape) == tuple(gshape):
        var_dict[cname] = gname
  return var_dict


def global_variables(params, with_ema=True):
  """[(name, shape)] of tf.global_variables() of the TRAIN graph (model variables, EMA shadows, global_step,
  Momentum slots) - the candidates `replace_initializers` walks."""
  mv = model_variables(params)
  out = list(mv)
  if with_ema:
    out += [(ema_name(n), s) for n, s in mv if has_ema(n)]
  out.append((GLOBAL_STEP, ()))
  out += [(momentum_name(n), s) for n, s in mv if trainable(n)]
  return out


def warm_start(params, ckpt_variables, psp_module=False):
  """Initialise from --init_ckpt_path (an exported ImageNet resnet_v1_50 checkpoint): matched variables take
  the checkpoint's values, everything else keeps its fresh initial value.  Returns the mapping used."""
  ckpt_vars = [(k, tuple(v.shape)) for k, v in ckpt_variables.items()]
  mapping = match_init_checkpoint(ckpt_vars, global_variables(params), psp_module)
  current = params.to_tf_dict()
  for cname, gname in mapping.items():
    current[gname] = torch.as_tensor(ckpt_variables[cname]).to(torch.float32)
  params.load_tf_dict(current)
  return mapping
