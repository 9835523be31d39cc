"""Golden vectors of predict.py and evaluate.py produced by running the reference's own scripts:
tests/golden/reference_predict_run.npz.

    python tests/golden/make_reference_predict_fixtures.py       # needs /root/reference; run in the build container

`code/predict.py::main(argv)` is executed UNMODIFIED with tests/golden/tf_shim first on sys.path: its argument parser
(SemanticSegmentationArguments PREDICT + add_predict_input_pipeline_arguments + add_model_arguments + the dataset
positional + _add_predict_arguments), `_add_extra_args`, `_validate_settings`, the reference's `SemanticSegmentation.
__init__ / .predict()` and the export loop (predict.py:137-164: label-id PNG through cids2lids, colour PNG through
cids2colors, 50:50 overlay on the raw image, file names from `split_path(str(rawimagespaths))`).  tf.estimator.Estimator
is a stub whose predict() yields two fixed examples (decisions incl. the void class, raw images, byte paths) and records
what it was asked for; matplotlib (absent here, imported at the top of predict.py, used only by the live-plotting flags)
is an empty module.  Stored: the parsed settings, predict_keys, checkpoint path handed to the estimator, the file names
written and the decoded PNGs.  tests/test_reference_fixtures.py runs wlseg.settings + wlseg.cli.export_outputs on the
same examples and compares names and pixels.

`code/evaluate.py::main(argv)` (the `__main__` guard raises upstream, main() itself is intact) runs the same way over the
stub estimator of make_reference_driver_fixtures.py (a fixed 20 x 20 confusion matrix, global_step 1234): stored are the
text of `eval_00/all_metrics.txt` (print_metrics_from_confusion_matrix through `printfile`) and the unpickled
`all_metrics.p`.  The test runs wlseg.cli.evaluate_main on the same argv with the same fake estimator.
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
OUT = os.path.join(HERE, 'reference_predict_run.npz')
SEED = 53
ARGV_TAIL = ['problem_definitions/cityscapes/problem01.json', 'some/predict/dir', 'cityscapes', '--export_lids_images',
             '--export_color_decisions', '--export_overlapped_color_decisions', '--ckpt_path', 'model.ckpt-12631']


EVAL_ARGV_TAIL = ['500', 'problem_definitions/cityscapes/problem01.json', 'tfrecords/x.tfrecords', 'cityscapes']


def examples():
  """Two per-image dicts as tf.estimator's predict() yields them for the reference's predict_keys."""
  rng = np.random.default_rng(SEED)
  out = []
  for i, (h, w, path) in enumerate(((12, 20, b'/data/demo/frankfurt_000000_000294_leftImg8bit.png'), (9, 14, b'relative/dir/img.0001.jpg'))):
    decs = rng.integers(0, 20, size=(h, w)).astype(np.int32)      # 19 = the void class of the training definition
    out.append({'decisions': decs, 'l1_probabilities': rng.random((h, w, 14)).astype(np.float32),
                'l2_vehicle_probabilities': rng.random((h, w, 7)).astype(np.float32),
                'rawimages': rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8), 'rawimagespaths': path})
  return out


def main():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import importlib
  import tensorflow as tf
  from make_reference_driver_fixtures import install_stubs, jsonable
  importlib.import_module('tensorflow.gfile')
  importlib.import_module('tensorflow.contrib.distribute')
  calls = []
  install_stubs(tf, calls)
  estimator_cls = tf.estimator.Estimator

  def predict(self, input_fn=None, predict_keys=None, checkpoint_path=None, **kw):
    calls.append(('predict', {'predict_keys': list(predict_keys), 'checkpoint_path': checkpoint_path}))
    for ex in examples():
      yield {k: ex[k] for k in predict_keys}
  estimator_cls.predict = predict
  plt = types.ModuleType('matplotlib.pyplot')
  plt.ion = lambda: None
  sys.modules['matplotlib'] = types.ModuleType('matplotlib')
  sys.modules['matplotlib.pyplot'] = plt
  sys.modules['matplotlib'].pyplot = plt

  cwd = os.getcwd()
  os.chdir(REF)          # the reference reads its problem definitions by relative path (read only)
  out = {}
  try:
    import predict as rpredict
    from PIL import Image
    captured = {}
    reference_system = rpredict.SemanticSegmentation

    def recording_system(*a, **k):       # the reference's class; the instance is kept to read its settings afterwards
      captured['system'] = reference_system(*a, **k)
      return captured['system']
    rpredict.SemanticSegmentation = recording_system
    with tempfile.TemporaryDirectory() as tmp:
      results = os.path.join(tmp, 'results')
      os.makedirs(results)
      with contextlib.redirect_stdout(io.StringIO()):
        rpredict.main([os.path.join(tmp, 'log')] + ARGV_TAIL + ['--results_dir', results])
      names = sorted(os.listdir(results))
      out['files'] = np.asarray('\n'.join(names))
      for n in names:
        out[f'png/{n}'] = np.asarray(Image.open(os.path.join(results, n)))
    st = captured['system'].settings
    out['settings'] = np.asarray(json.dumps({k: v for k, v in vars(st).items() if jsonable(v) and k not in (
        'log_dir', 'results_dir', 'eval_res_dir', 'training_problem_def', 'inference_problem_def', 'evaluation_problem_def')}, sort_keys=True))
    out['calls'] = np.asarray(json.dumps([c for c in calls if c[0] == 'predict']))
    out['argv'] = np.asarray(json.dumps(ARGV_TAIL))
    # ---- evaluate.py::main
    import pickle
    import evaluate as revaluate
    del calls[:]
    with tempfile.TemporaryDirectory() as tmp:
      log_dir = os.path.join(tmp, 'log')
      os.makedirs(log_dir)
      with contextlib.redirect_stdout(io.StringIO()):
        revaluate.main([log_dir] + EVAL_ARGV_TAIL)
      res = os.path.join(log_dir, 'eval_00')
      out['evaluate/files'] = np.asarray('\n'.join(sorted(os.listdir(res))))
      with open(os.path.join(res, 'all_metrics.txt')) as fp:
        out['evaluate/all_metrics_txt'] = np.asarray(fp.read())
      with open(os.path.join(res, 'all_metrics.p'), 'rb') as fp:
        pickled = pickle.load(fp)
      out['evaluate/pickle_keys'] = np.asarray('\n'.join(sorted(pickled[0].keys())))
      out['evaluate/pickle_len'] = np.asarray(len(pickled))
      out['evaluate/pickle_cm'] = np.asarray(pickled[0]['confusion_matrix'])
      out['evaluate/pickle_global_step'] = np.asarray(pickled[0]['global_step'])
    out['evaluate/argv'] = np.asarray(json.dumps(EVAL_ARGV_TAIL))
    for i, ex in enumerate(examples()):
      out[f'example{i}/decisions'] = ex['decisions']
      out[f'example{i}/rawimages'] = ex['rawimages']
      out[f'example{i}/rawimagespaths'] = np.asarray(ex['rawimagespaths'].decode())
  finally:
    os.chdir(cwd)
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes;', str(out['files']).split('\n'))


if __name__ == '__main__':
  sys.path.insert(0, HERE)
  main()
