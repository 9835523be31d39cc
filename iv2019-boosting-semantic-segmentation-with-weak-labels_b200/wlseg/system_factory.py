"""`SemanticSegmentation`: the reference's system facade on the B200 kernels.

Mirrors code/system_factory.py:27-461: same constructor signature
`SemanticSegmentation(input_fns, model_fn, settings)`, same derived settings
(`output_Nclasses` :124-130, cid maps :138-157, steps and LR schedule :197-233, checkpoint cadence
:246-248, eval directory naming :159-172), same `.train()`, `.evaluate() -> list of metrics dicts`
(void row / column trimmed, :400-405), `.predict() -> iterator of per-image dicts`, `.settings`.
`tf.estimator.Estimator` is replaced by wlseg.estimator.Estimator.
"""

import collections
import copy
import glob
import json
import os
from os.path import exists, isdir, join, split

from wlseg import estimator as est
from wlseg import hierarchy, metrics


class SemanticSegmentation(object):

  def __init__(self, input_fns, model_fn, settings):
    assert settings is not None, ('settings must be provided for now.')
    self._settings = copy.deepcopy(settings)
    s = self._settings
    s.height_network = s.height_feature_extractor
    s.width_network = s.width_feature_extractor

    with open(s.training_problem_def_path, 'r') as fp:
      s.training_problem_def = json.load(fp)
    if hasattr(s, 'inference_problem_def_path'):
      if s.inference_problem_def_path:
        with open(s.inference_problem_def_path, 'r') as fp:
          s.inference_problem_def = json.load(fp)
      else:
        s.inference_problem_def = s.training_problem_def
    if hasattr(s, 'evaluation_problem_def_path'):
      if s.evaluation_problem_def_path:
        with open(s.evaluation_problem_def_path, 'r') as fp:
          s.evaluation_problem_def = json.load(fp)
      else:
        s.evaluation_problem_def = s.training_problem_def

    _set_defaults(s)
    _validate_settings(s)

    self._input_fns = input_fns
    self._model_fn = model_fn
    self._estimator = None

    lids2cids_training = s.training_problem_def['lids2cids']
    s.lids_training_contain_unlabeled = -1 in lids2cids_training
    s.output_Nclasses = (max(lids2cids_training) + 1 +
                         (s.lids_training_contain_unlabeled or s.train_void_class))

    if hasattr(s, 'inference_problem_def'):
      if 'training_cids2inference_cids' in s.inference_problem_def.keys():
        s.training_cids2inference_cids = s.inference_problem_def['training_cids2inference_cids']
      else:
        tcids2pcids = list(range(s.output_Nclasses))
        if s.lids_training_contain_unlabeled and not s.train_void_class:
          tcids2pcids[-1] = -1
        s.training_cids2inference_cids = tcids2pcids
    if hasattr(s, 'evaluation_problem_def'):
      if 'training_cids2evaluation_cids' in s.evaluation_problem_def.keys():
        s.training_cids2evaluation_cids = s.evaluation_problem_def['training_cids2evaluation_cids']
      else:
        tcids2ecids = list(range(s.output_Nclasses))
        if s.lids_training_contain_unlabeled and not s.train_void_class:
          tcids2ecids[-1] = -1
        s.training_cids2evaluation_cids = tcids2ecids

    existing_eval_dirs = list(filter(isdir, glob.glob(join(s.log_dir, 'eval_*'))))
    if existing_eval_dirs:
      max_cnt = max([int(split(ed)[1][-2:]) for ed in existing_eval_dirs])
    else:
      max_cnt = -1
    s.eval_res_dir = join(s.log_dir, 'eval_' + f"{max_cnt + 1:02}")

    # the hierarchy tables are not part of the problem definition upstream (hard-coded per dataset);
    # here they are derived from its class names
    self._hier = hierarchy.Hierarchy(s.per_pixel_dataset_name, s.training_problem_def['cids2labels'])
    assert self._hier.num_classes == s.output_Nclasses, (
        f"problem definition has {s.output_Nclasses} classes but the {s.per_pixel_dataset_name} "
        f"hierarchy expects {self._hier.num_classes}")

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
    s.num_examples_per_epoch = int(s.Ntrain * s.height_network // s.height_feature_extractor *
                                   s.width_network // s.width_feature_extractor)
    s.num_batches_per_epoch = int(s.num_examples_per_epoch / s.Nb)
    s.num_training_steps = int(s.Ne * s.num_batches_per_epoch)

    if s.learning_rate_schedule == 'piecewise_constant':
      if not (s.learning_rate_decay or s.learning_rate_values):
        s.learning_rate_decay = 0.5
      last_boundary = s.Ne - s.learning_rate_boundaries[-1]
      if last_boundary == 0:
        s.learning_rate_boundaries.pop()
      elif last_boundary < 0:
        raise ValueError('Ne is less than learning rate boundaries.')
      s.learning_rate_boundaries_epochs = s.learning_rate_boundaries
      s.learning_rate_boundaries = [lrb * s.num_batches_per_epoch for lrb in s.learning_rate_boundaries]
      if s.learning_rate_decay:
        decay_steps = len(s.learning_rate_boundaries) + 1
        s.learning_rate_values = [s.learning_rate_initial * s.learning_rate_decay ** i for i in range(decay_steps)]

    if s.distribute:
      print('\n\nDisabling moving running averages for distributed training.\n\n')
      s.ema_decay = 0

    os.makedirs(s.log_dir, exist_ok=True)
    if not s.save_checkpoints_steps:
      s.save_checkpoints_steps = s.num_batches_per_epoch

    settings_dict = collections.OrderedDict(sorted(vars(s).items()))
    settings_filename = join(s.log_dir, 'settings.txt')
    # one process per GPU: rank 0 owns the log directory (the reference is a single process); its verdict is made
    # collective so that a failed precondition stops EVERY rank instead of leaving the others in their first collective
    stale = bool(getattr(s, 'rank', 0) == 0 and exists(settings_filename))
    stale = _any_rank(stale, s)
    assert not stale, (
        f"Previous settings.txt found in {s.log_dir}. Rename or delete it manually and restart training.")
    if getattr(s, 'rank', 0) == 0:
      with open(settings_filename, 'w') as f:
        for k, v in enumerate(settings_dict):
          print(f"{k:2} : {v} : {settings_dict[v]}", file=f)

    self._create_estimator(for_training=True)
    max_steps = s.num_training_steps if not getattr(s, 'steps', None) else min(s.steps, s.num_training_steps)
    return self._estimator.train(self._input_fns['train'](None, s), max_steps)

  # ------------------------------------------------------------------------------------------ predict
  def predict(self):
    s = self._settings
    if s.Nb > 1:
      print('\nWARNING: during prediction only images with same shape (size and channels) '
            'are supported for batch size greater than one. In case of runtime error '
            'change batch size to 1.\n')
    self._create_estimator(ckpt_path=s.ckpt_path)
    predict_keys = copy.deepcopy(s.predict_keys)
    return self._estimator.predict(self._input_fns['predict'](None, s), predict_keys)

  # ------------------------------------------------------------------------------------------ evaluate
  def evaluate(self):
    s = self._settings
    s.num_examples = int(s.Neval * s.height_network // s.height_feature_extractor *
                         s.width_network // s.width_feature_extractor)
    s.num_batches_per_epoch = int(s.num_examples / s.Nb)
    s.num_eval_steps = int(s.num_batches_per_epoch * 1)

    eval_res_dir = s.eval_res_dir
    print(f"\nWriting results in {eval_res_dir}.\n")
    if getattr(s, 'rank', 0) == 0:
      os.makedirs(eval_res_dir)
      if exists(join(eval_res_dir, 'settings.txt')):
        print(f"WARNING: previous settings.txt in {eval_res_dir} is ovewritten.")
      with open(join(eval_res_dir, 'settings.txt'), 'w') as f:
        for k, v in vars(s).items():
          print(f"{k} : {v}", file=f)

    labels = s.evaluation_problem_def['cids2labels']
    void_exists = -1 in s.evaluation_problem_def['lids2cids']
    if void_exists and not s.train_void_class:
      labels = labels[:-1]

    if getattr(s, 'preserve_aspect_ratio', False):
      raise NotImplementedError('evaluation with preserving aspect ratio is not implemented.')

    all_model_checkpoint_paths = [s.ckpt_path]
    if s.eval_all_ckpts:
      s.ckpt_path = None
      all_model_checkpoint_paths = sorted(glob.glob(join(s.log_dir, 'model.ckpt-*.pt')),
                                          key=lambda p: int(p.rsplit('-', 1)[1].split('.')[0]))
      print(f"\n{len(all_model_checkpoint_paths)} checkpoint(s) will be evaluated.\n")

    tcids2ecids = est._replacevoids(s.training_cids2evaluation_cids)
    num_classes = max(tcids2ecids) + 1
    identity = tcids2ecids == list(range(len(tcids2ecids)))

    all_metrics = []
    for cp in all_model_checkpoint_paths:
      self._create_estimator(ckpt_path=cp)
      metrics_ = self._estimator.evaluate(self._input_fns['eval'](None, s), num_classes,
                                          lut=None if identity else tcids2ecids)
      metrics_ = self._reduce_across_ranks(metrics_)
      if (-1 in s.evaluation_problem_def['lids2cids'] and not s.train_void_class):
        metrics_['confusion_matrix'] = metrics_['confusion_matrix'][:-1, :-1]
      if getattr(s, 'rank', 0) == 0:
        metrics.print_metrics_from_confusion_matrix(metrics_['confusion_matrix'], labels, printcmd=True)
      all_metrics.append(metrics_)
    return all_metrics

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


def _set_defaults(settings):
  if hasattr(settings, 'learning_rate_schedule'):
    if settings.learning_rate_schedule == 'piecewise_constant':
      if not (settings.learning_rate_decay or settings.learning_rate_values):
        settings.learning_rate_decay = 0.5


def _validate_settings(settings):
  assert all([settings.height_network == settings.height_feature_extractor,
              settings.width_network == settings.width_feature_extractor]), (
                  f"For now height_network ({settings.height_network}), "
                  f"height feature_extractor ({settings.height_feature_extractor}), "
                  f"and width_network ({settings.width_network}), "
                  f"width_feature_extractor ({settings.width_feature_extractor}) "
                  "should be equal.")
  if hasattr(settings, 'learning_rate_schedule'):
    if settings.learning_rate_schedule == 'piecewise_constant':
      if not (bool(settings.learning_rate_decay) != bool(settings.learning_rate_values)):
        raise AttributeError('If `learning_rate_schedule` is `piecewise_constant` exactly one of '
                             '`learning_rate_decay` or `learning_rate_values` must be given.')
  lids2cids_unique = set(settings.training_problem_def['lids2cids'])
  cid_max = max(lids2cids_unique)
  lids2cids_unique.discard(-1)
  if not (lids2cids_unique == set(range(cid_max + 1))):
    raise ValueError('lids2cids field in training problem definition contains not continuous class ids.')
