"""`SemanticSegmentation`: the reference's system facade on the B200 kernels.

Drop-in for code/system_factory.py:27-461: the constructor `SemanticSegmentation(input_fns, model_fn, settings)`,
`.train()`, `.evaluate() -> [metrics dict per checkpoint]`, `.predict() -> iterator of per-image dicts`, `.settings`.
Every attribute the reference leaves on `settings` is derived here as well (tests/golden/reference_driver_run.json holds
the reference's own values for 56 training and 37 evaluation attributes; tests/test_reference_fixtures.py compares):

  derivation (pure functions below)                 reference
  ------------------------------------------------  ----------------------------
  problem definitions, fall-backs                   system_factory.py:95-117
  class count, training -> inference / eval maps    :124-157
  eval_NN result directory                          :159-172
  steps per epoch / total steps                     :197-201, :355-361
  piecewise schedule: epochs -> steps, plateau lrs  :207-233
  checkpoint cadence, EMA switch under --distribute :238-248
  void row / column of the confusion matrix         :400-405

`tf.estimator.Estimator` is replaced by wlseg.estimator.Estimator; one process per GPU replaces MirroredStrategy, so the
log-directory preconditions are decided on rank 0 and agreed on collectively.
"""

import copy
import glob
import json
import os

from wlseg import estimator as est
from wlseg import hierarchy, metrics

_VOID = -1


# ---------------------------------------------------------------------------------------------- pure derivations
def _read_json(path):
  with open(path, 'r') as fp:
    return json.load(fp)


def attach_problem_definitions(s):
  """training_problem_def always; inference / evaluation ones when the mode has the flag, falling back to the training
  definition when the flag is empty."""
  s.training_problem_def = _read_json(s.training_problem_def_path)
  for kind in ('inference', 'evaluation'):
    flag = f'{kind}_problem_def_path'
    if hasattr(s, flag):
      path = getattr(s, flag)
      setattr(s, f'{kind}_problem_def', _read_json(path) if path else s.training_problem_def)


def check_settings(s):
  """The reference's `_validate_settings` (:431-461): network size == feature-extractor size, exactly one of
  decay / values for the piecewise schedule, contiguous training class ids."""
  if (s.height_network, s.width_network) != (s.height_feature_extractor, s.width_feature_extractor):
    raise AssertionError(f'network size {s.height_network}x{s.width_network} must equal the feature extractor size '
                         f'{s.height_feature_extractor}x{s.width_feature_extractor} for now.')
  if getattr(s, 'learning_rate_schedule', None) == 'piecewise_constant':
    if bool(s.learning_rate_decay) == bool(s.learning_rate_values):
      raise AttributeError('piecewise_constant needs exactly one of learning_rate_decay / learning_rate_values.')
  cids = set(s.training_problem_def['lids2cids']) - {_VOID}
  if cids != set(range(max(cids) + 1)):
    raise ValueError('lids2cids of the training problem definition must cover 0..max without gaps.')


def class_id_space(s):
  """-> (has unlabeled ids, number of output classes): one extra class when label ids map to void or a void class is
  trained explicitly."""
  lids2cids = s.training_problem_def['lids2cids']
  unlabeled = _VOID in lids2cids
  return unlabeled, max(lids2cids) + 1 + (unlabeled or s.train_void_class)


def training_cids_to(problem_def, key, n_classes, last_is_void):
  """Map from training class ids to the ids of another problem definition: the definition's own table when it has one,
  else the identity with the trailing void class sent to -1."""
  if key in problem_def:
    return problem_def[key]
  ids = list(range(n_classes))
  if last_is_void:
    ids[-1] = _VOID
  return ids


def next_eval_dir(log_dir):
  """log_dir/eval_NN with NN one above the highest existing."""
  taken = [int(os.path.basename(d)[-2:]) for d in glob.glob(os.path.join(log_dir, 'eval_*')) if os.path.isdir(d)]
  return os.path.join(log_dir, f'eval_{max(taken, default=-1) + 1:02}')


def steps_per_epoch(n_examples, s):
  """(examples per epoch, batches per epoch); the size ratio is 1 while network == feature-extractor size."""
  examples = int(n_examples * s.height_network // s.height_feature_extractor * s.width_network // s.width_feature_extractor)
  return examples, int(examples / s.Nb)


def piecewise_schedule_in_steps(s):
  """Boundaries given in epochs become steps; a boundary at the last epoch is dropped; with a decay factor the plateau
  values are lr0 * decay^i."""
  if not (s.learning_rate_decay or s.learning_rate_values):
    s.learning_rate_decay = 0.5
  epochs = s.learning_rate_boundaries
  if epochs[-1] > s.Ne:
    raise ValueError('Ne is less than learning rate boundaries.')
  if epochs[-1] == s.Ne:
    epochs.pop()
  s.learning_rate_boundaries_epochs = epochs
  s.learning_rate_boundaries = [e * s.num_batches_per_epoch for e in epochs]
  if s.learning_rate_decay:
    s.learning_rate_values = [s.learning_rate_initial * s.learning_rate_decay ** i for i in range(len(epochs) + 1)]


def _rank(s):
  return getattr(s, 'rank', 0)


def _any_rank(flag, settings):
  """Logical OR of `flag` over the ranks (identity in a single process)."""
  import torch
  import torch.distributed as dist
  if getattr(settings, 'world_size', 1) > 1 and dist.is_available() and dist.is_initialized():
    dev = getattr(settings, 'device', 'cuda') if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))
  return bool(flag)


# ---------------------------------------------------------------------------------------------- the facade
class SemanticSegmentation(object):

  def __init__(self, input_fns, model_fn, settings):
    if settings is None:
      raise AssertionError('settings must be provided for now.')
    s = self._settings = copy.deepcopy(settings)
    self._input_fns, self._model_fn, self._estimator = input_fns, model_fn, None

    s.height_network, s.width_network = s.height_feature_extractor, s.width_feature_extractor
    attach_problem_definitions(s)
    if getattr(s, 'learning_rate_schedule', None) == 'piecewise_constant' and not (s.learning_rate_decay or s.learning_rate_values):
      s.learning_rate_decay = 0.5
    check_settings(s)

    s.lids_training_contain_unlabeled, s.output_Nclasses = class_id_space(s)
    last_is_void = s.lids_training_contain_unlabeled and not s.train_void_class
    for kind in ('inference', 'evaluation'):
      if hasattr(s, f'{kind}_problem_def'):
        key = f'training_cids2{kind}_cids'
        setattr(s, key, training_cids_to(getattr(s, f'{kind}_problem_def'), key, s.output_Nclasses, last_is_void))
    s.eval_res_dir = next_eval_dir(s.log_dir)

    # upstream hard-codes the hierarchy tables per dataset; here they come from the class names of the definition
    self._hier = hierarchy.Hierarchy(s.per_pixel_dataset_name, s.training_problem_def['cids2labels'])
    if self._hier.num_classes != s.output_Nclasses:
      raise AssertionError(f'problem definition has {s.output_Nclasses} classes but the {s.per_pixel_dataset_name} '
                           f'hierarchy expects {self._hier.num_classes}')

  @property
  def settings(self):
    return self._settings

  @property
  def estimator(self):
    return self._estimator

  def _create_estimator(self, ckpt_path=None, for_training=False):
    s = self._settings
    if getattr(s, 'name_feature_extractor', 'resnet_v1_50') == 'resnet_v1_101':
      # code/estimator/define_estimator_hierarchical.py:57-61
      raise NotImplementedError('Use of resnet_v1_101 as base feature extractor is not yet implemented.')
    if bool(getattr(s, 'fov_expansion_kernel_rate', 0)) != bool(getattr(s, 'fov_expansion_kernel_size', 0)):
      # _validate_params, code/models/resnet50_extended_model_hierarchical.py:271-276
      raise ValueError('One of params.{fov_expansion_kernel_rate, fov_expansion_kernel_size} '
                       'is set. In order to take effect both should be set.')
    if getattr(s, 'norm_layer', 'batch') == 'group' and getattr(s, 'cross_replica_norm', False):
      # module_arg_scope, code/models/resnet50_extended_model_hierarchical.py:331-333
      raise ValueError('cross_replica_norm is supported only for batch normalization for now.')
    if for_training and getattr(s, 'upsampling_method', 'bilinear') == 'no':
      # the reference's graph fails here too: per-pixel labels (hf x wf) against logits at hf/8 x wf/8
      raise ValueError('--upsampling_method no: labels and logits differ in size, training is not possible.')
    self._estimator = est.Estimator(s, self._hier, device=getattr(s, 'device', 'cuda'))
    self._estimator.initialize(ckpt_path=ckpt_path, log_dir=s.log_dir, seed=getattr(s, 'seed', 0),
                               for_training=for_training)
    return self._estimator

  # ------------------------------------------------------------------------------------------ train
  def train(self):
    s = self._settings
    s.num_examples_per_epoch, s.num_batches_per_epoch = steps_per_epoch(s.Ntrain, s)
    s.num_training_steps = int(s.Ne * s.num_batches_per_epoch)
    if s.learning_rate_schedule == 'piecewise_constant':
      piecewise_schedule_in_steps(s)
    if s.distribute:
      print('\n--distribute: exponential moving averages are switched off.\n')
      s.ema_decay = 0
    os.makedirs(s.log_dir, exist_ok=True)
    s.save_checkpoints_steps = s.save_checkpoints_steps or s.num_batches_per_epoch

    # settings.txt: one "index : name : value" line per attribute, sorted by name; a previous file stops the run.
    # Rank 0 owns the log directory (the reference is a single process) and every rank learns its verdict, so that a
    # failed precondition stops the whole job instead of leaving the other ranks in their first collective.
    record = os.path.join(s.log_dir, 'settings.txt')
    if _any_rank(_rank(s) == 0 and os.path.exists(record), s):
      raise AssertionError(f'Previous settings.txt found in {s.log_dir}. Rename or delete it manually and restart training.')
    if _rank(s) == 0:
      with open(record, 'w') as fp:
        for i, (name, value) in enumerate(sorted(vars(s).items())):
          print(f'{i:2} : {name} : {value}', file=fp)

    self._create_estimator(for_training=True)
    budget = getattr(s, 'steps', None)
    max_steps = min(budget, s.num_training_steps) if budget else s.num_training_steps
    return self._estimator.train(self._input_fns['train'](None, s), max_steps)

  # ------------------------------------------------------------------------------------------ predict
  def predict(self):
    s = self._settings
    if s.Nb > 1:
      print('\nWARNING: a prediction batch of more than one image needs images of one shape; use --Nb 1 otherwise.\n')
    self._create_estimator(ckpt_path=s.ckpt_path)
    return self._estimator.predict(self._input_fns['predict'](None, s), list(s.predict_keys))

  # ------------------------------------------------------------------------------------------ evaluate
  def evaluate(self):
    s = self._settings
    s.num_examples, s.num_batches_per_epoch = steps_per_epoch(s.Neval, s)
    s.num_eval_steps = int(s.num_batches_per_epoch * 1)
    if getattr(s, 'preserve_aspect_ratio', False):
      raise NotImplementedError('evaluation with preserving aspect ratio is not implemented.')

    print(f'\nWriting results in {s.eval_res_dir}.\n')
    if _rank(s) == 0:
      os.makedirs(s.eval_res_dir)
      with open(os.path.join(s.eval_res_dir, 'settings.txt'), 'w') as fp:
        for name, value in vars(s).items():
          print(f'{name} : {value}', file=fp)

    # the void class (last id) is evaluated only when it was trained explicitly
    drop_void = _VOID in s.evaluation_problem_def['lids2cids'] and not s.train_void_class
    class_names = s.evaluation_problem_def['cids2labels']
    if drop_void:
      class_names = class_names[:-1]

    checkpoints = [s.ckpt_path]
    if s.eval_all_ckpts:
      s.ckpt_path = None
      checkpoints = sorted(glob.glob(os.path.join(s.log_dir, 'model.ckpt-*.pt')),
                           key=lambda p: int(p.rsplit('-', 1)[1].split('.')[0]))
      print(f'\n{len(checkpoints)} checkpoint(s) will be evaluated.\n')

    lut = est._replacevoids(s.training_cids2evaluation_cids)
    num_classes = max(lut) + 1
    if lut == list(range(len(lut))):
      lut = None                      # identity map: the kernel skips the lookup

    results = []
    for path in checkpoints:
      self._create_estimator(ckpt_path=path)
      m = self._reduce_across_ranks(self._estimator.evaluate(self._input_fns['eval'](None, s), num_classes, lut=lut))
      if drop_void:
        m['confusion_matrix'] = m['confusion_matrix'][:-1, :-1]
      if _rank(s) == 0:
        metrics.print_metrics_from_confusion_matrix(m['confusion_matrix'], class_names, printcmd=True)
      results.append(m)
    return results

  def _reduce_across_ranks(self, m):
    """Evaluation sharded by image: integer confusion matrices are summed across ranks (exact)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if getattr(self._settings, 'world_size', 1) > 1 and dist.is_available() and dist.is_initialized():
      dev = self._estimator.device if dist.get_backend() == 'nccl' else 'cpu'
      t = torch.from_numpy(m['confusion_matrix_int64']).to(dev)
      dist.all_reduce(t, op=dist.ReduceOp.SUM)
      m['confusion_matrix_int64'] = t.cpu().numpy()
      m['confusion_matrix'] = m['confusion_matrix_int64'].astype(np.int32)
    return m
