"""Golden vectors of the DRIVER layer produced by running the reference's own Python:  tests/golden/reference_driver_run.json.

    python tests/golden/make_reference_driver_fixtures.py       # needs /root/reference; run in the build container

Executed, unmodified, with tests/golden/tf_shim first on sys.path (tf.estimator.Estimator / RunConfig / ConfigProto are
recording stubs added here - nothing is trained):
  utils.utils.SemanticSegmentationArguments + add_train_input_pipeline_arguments + add_model_arguments   (the CLI, with defaults)
  train._add_extra_args                                                                                  (train.py:42-68)
  system_factory.SemanticSegmentation.__init__ / .train() / .evaluate()                                  (system_factory.py:52-412)
Stored per case: the argv, every JSON-representable attribute of `system.settings` after the call (derived class counts, id
maps, steps per epoch, total steps, learning-rate boundaries in steps and values, checkpoint cadence, EMA switch, evaluation
steps ...), what the Estimator stub was asked to do, and for evaluate() the confusion matrix it returns for a given raw one
(void row / column trimmed).  tests/test_reference_fixtures.py runs wlseg.settings + wlseg.system_factory on the same
argv and compares.
"""

import contextlib
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('WLSEG_REFERENCE', '/root/reference/code')
OUT = os.path.join(HERE, 'reference_driver_run.json')

TRAIN_CASES = {
    'train_cityscapes_defaults': ['cityscapes'],
    'train_vistas_defaults': ['vistas'],
    'train_cityscapes_void_poly': ['cityscapes', '--train_void_class', '--learning_rate_schedule', 'polynomial_decay', '--Ne', '5',
                                   '--Nb', '2', '--optimizer', 'SGD', '--use_nesterov', '--ema_decay', '0.99'],
    'train_cityscapes_flags': ['cityscapes', '--Ne', '30', '--learning_rate_initial', '0.02', '--learning_rate_boundaries', '10', '20', '30',
                               '--distribute', '--psp_module', '--save_checkpoints_steps', '500'],
}
EVAL_CASES = {
    'eval_cityscapes': (['500', 'problem_definitions/cityscapes/problem01.json', 'tfrecords/x.tfrecords', 'cityscapes'], 20),
    'eval_vistas': (['2000', 'problem_definitions/vistas/problem01.json', 'tfrecords/x.tfrecords', 'vistas', '--Nb', '2'], 66),
    # --train_void_class: the void class is a trained class - the id map keeps it and the matrix is handed back untrimmed
    'eval_cityscapes_train_void': (['500', 'problem_definitions/cityscapes/problem01.json', 'tfrecords/x.tfrecords', 'cityscapes',
                                    '--train_void_class', '--restore_emas', '--Nb', '4'], 20),
}


def jsonable(v):
  try:
    json.dumps(v)
    return True
  except (TypeError, ValueError):
    return False


def install_stubs(tf, calls):
  class Estimator:
    def __init__(self, model_fn, model_dir=None, config=None, params=None, **kw):
      calls.append(('Estimator', {'model_dir': model_dir}))
      self.params = params

    def train(self, input_fn=None, max_steps=None, steps=None, **kw):
      calls.append(('train', {'max_steps': max_steps, 'steps': steps}))
      return self

    def evaluate(self, input_fn=None, steps=None, checkpoint_path=None, name=None, **kw):
      calls.append(('evaluate', {'steps': steps, 'checkpoint_path': checkpoint_path, 'name': name}))
      C = self.params.output_Nclasses
      cm = (np.arange(C * C, dtype=np.int32).reshape(C, C) % 7) + np.eye(C, dtype=np.int32) * 50
      return {'global_step': 1234, 'loss': 0.5, 'confusion_matrix': cm}

  class RunConfig:
    def __init__(self, **kw):
      calls.append(('RunConfig', {k: v for k, v in kw.items() if jsonable(v)}))
      self.train_distribute = kw.get('train_distribute')

  def config_proto():
    return types.SimpleNamespace(gpu_options=types.SimpleNamespace(allow_growth=False), allow_soft_placement=False,
                                 graph_options=types.SimpleNamespace(optimizer_options=types.SimpleNamespace(global_jit_level=0)))
  tf.estimator.Estimator = Estimator
  tf.estimator.RunConfig = RunConfig
  tf.ConfigProto = config_proto
  tf.OptimizerOptions = types.SimpleNamespace(ON_1=1)
  # module-level tf.data feature tables of the input pipelines (import time only)
  tf.FixedLenFeature = lambda *a, **k: ('FixedLenFeature', a)
  tf.VarLenFeature = lambda *a, **k: ('VarLenFeature', a)
  tf.string = 'string'
  sys.modules['tensorflow.train'].latest_checkpoint = lambda d: None     # fresh log directories
  sys.modules['tensorflow.gfile'].Exists = os.path.exists
  sys.modules['tensorflow.gfile'].MakeDirs = os.makedirs
  sys.modules['tensorflow.contrib.distribute'].MirroredStrategy = lambda *a, **k: 'MirroredStrategy'


def main():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import tensorflow as tf
  import importlib
  importlib.import_module('tensorflow.gfile')
  importlib.import_module('tensorflow.contrib.distribute')
  calls = []
  install_stubs(tf, calls)
  cwd = os.getcwd()
  os.chdir(REF)          # the reference reads its problem definitions by relative path (read only)
  out = {}
  try:
    import train as rtrain
    from system_factory import SemanticSegmentation
    from utils.utils import SemanticSegmentationArguments
    from models.resnet50_extended_model_hierarchical import add_model_arguments
    from input_pipelines.heterogeneous_supervision.per_pixel_per_bbox_per_image import add_train_input_pipeline_arguments
    for tag, argv in TRAIN_CASES.items():
      del calls[:]
      with tempfile.TemporaryDirectory() as log_dir:
        ssargs = SemanticSegmentationArguments(mode=tf.estimator.ModeKeys.TRAIN)
        add_train_input_pipeline_arguments(ssargs.argparser)
        add_model_arguments(ssargs.argparser)
        settings = ssargs.parse_args([log_dir] + argv)
        parsed = {k: v for k, v in vars(settings).items() if jsonable(v) and k != 'log_dir'}
        rtrain._add_extra_args(settings)
        system = SemanticSegmentation({'train': None}, None, settings)
        with contextlib.redirect_stdout(io.StringIO()):
          system.train()
        st = {k: v for k, v in vars(system.settings).items()
              if jsonable(v) and k not in ('log_dir', 'eval_res_dir', 'training_problem_def', 'inference_problem_def',
                                           'evaluation_problem_def')}
        out[tag] = {'argv': argv, 'parsed': parsed, 'settings': st,
                    'calls': [(n, {k: v for k, v in kw.items() if k != 'model_dir'}) for n, kw in calls]}
    import evaluate as revaluate
    from input_pipelines.cityscapes.input_cityscapes import add_evaluate_input_pipeline_arguments
    for tag, (argv, _) in EVAL_CASES.items():
      del calls[:]
      with tempfile.TemporaryDirectory() as log_dir:
        # evaluate.py:25-36 (main is disabled upstream by the raise at :82; the argument flow is its own)
        ssargs = SemanticSegmentationArguments(mode=tf.estimator.ModeKeys.EVAL)
        add_evaluate_input_pipeline_arguments(ssargs.argparser)
        add_model_arguments(ssargs.argparser)
        ssargs.argparser.add_argument('per_pixel_dataset_name', type=str, choices=['vistas', 'cityscapes'])
        settings = ssargs.parse_args([log_dir] + argv)
        parsed = {k: v for k, v in vars(settings).items() if jsonable(v) and k != 'log_dir'}
        revaluate._add_extra_args(settings)
        system = SemanticSegmentation({'eval': None}, None, settings)
        with contextlib.redirect_stdout(io.StringIO()):
          metrics = system.evaluate()
        st = {k: v for k, v in vars(system.settings).items()
              if jsonable(v) and k not in ('log_dir', 'eval_res_dir', 'training_problem_def', 'inference_problem_def',
                                           'evaluation_problem_def')}
        out[tag] = {'argv': argv, 'parsed': parsed, 'settings': st,
                    'eval_res_dir_name': os.path.basename(system.settings.eval_res_dir),
                    'calls': [(n, {k: v for k, v in kw.items() if k != 'model_dir'}) for n, kw in calls],
                    'returned_cm_shape': list(metrics[0]['confusion_matrix'].shape),
                    'returned_cm_sum': int(metrics[0]['confusion_matrix'].sum())}
  finally:
    os.chdir(cwd)
  with open(OUT, 'w') as fp:
    json.dump(out, fp, indent=1, sort_keys=True)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes;', {k: len(v['settings']) for k, v in out.items()})


if __name__ == '__main__':
  main()
