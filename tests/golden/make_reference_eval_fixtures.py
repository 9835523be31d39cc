"""Golden vectors of the EVAL and PREDICT branches produced by running the reference's own `define_estimator`:
tests/golden/reference_eval_run.npz.

    python tests/golden/make_reference_eval_fixtures.py       # needs /root/reference; run in the build container

`code/estimator/define_estimator_hierarchical.py::define_estimator` is imported UNMODIFIED with tests/golden/tf_shim
first on sys.path and called with the reference's own `model()` as `model_fn`:
  EVAL (:160-201), once per batch of the evaluation loop: model() in inference mode, define_losses (EVAL: zeros),
      `_map_predictions_to_new_cids(predictions, training_cids2evaluation_cids)`, `_resize_predictions` to the size of
      labels['prolabels'], `_replacevoids`, `metrics_impl._streaming_confusion_matrix(labels, decisions, max + 1)` and
      eval_metric_ops = (to_int32(total_cm), update_op).  Two cases: labels at the network's size, and labels at twice
      the network's size (Cityscapes: 512x1024 network, 1024x2048 labels - the nearest-neighbour resize of :530-571).
  PREDICT (:204-237): the four supported keys, `_resize_predictions` to (height_system, width_system), and - with
      either unset - to the size of features['rawimages'].
Restated (TF is un-vendored): tensorflow/_train.py::_streaming_confusion_matrix (float64 accumulator, int64 casts) next
to the shim's resize / softmax / argmax of the earlier fixtures.
Stored per case: the images and labels, per batch the remapped + resized decisions, the streaming confusion matrix after
the last batch (int32, as eval_metric_ops exposes it), the loss; for PREDICT the decisions and strided probabilities.
The parameters are `oracle.network.init_params` numbers (random only), rebuilt by the tests with the same call.
tests/test_reference_fixtures.py replays the run with the oracle on CPU; tests/test_gpu_reference_fixtures.py runs the
product's `Estimator.evaluate` / `.predict` on the same batches.
"""

import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('WLSEG_REFERENCE', '/root/reference/code')
OUT = os.path.join(HERE, 'reference_eval_run.npz')

SEED = 31
CITYSCAPES_TRAINING_CIDS2EVALUATION_CIDS = list(range(19)) + [-1]   # 19 evaluated classes + void (problem definition)
VISTAS_TRAINING_CIDS2EVALUATION_CIDS = list(range(65)) + [-1]
# tag -> (dataset, batches, N, network H, W, label H, W)
EVAL_CASES = {
    'eval_cs_same_size': ('cityscapes', 2, 2, 40, 56, 40, 56),
    'eval_cs_labels_2x': ('cityscapes', 2, 1, 40, 56, 80, 112),
    'eval_vistas_labels_odd': ('vistas', 1, 1, 40, 56, 53, 75),
    # --upsampling_method no: predictions stay at H/8 x W/8 and `_resize_predictions` carries them to the label size
    'eval_cs_no_upsampling': ('cityscapes', 1, 2, 40, 56, 40, 56),
}
# tag -> (dataset, N, network H, W, (height_system, width_system), raw image size or None)
PREDICT_CASES = {
    'predict_cs_system_size': ('cityscapes', 2, 40, 56, (64, 96), None),
    'predict_cs_raw_size': ('cityscapes', 1, 40, 56, (None, None), (50, 70)),
}
PROB_STRIDE = 3
PROB_KEYS = ('l1_probabilities', 'l2_vehicle_probabilities', 'l2_human_probabilities')


def case_params(dataset):
  if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
  from oracle import network as onet
  return onet.init_params(dataset, seed=SEED, randomize_bn=True, tame=True)


def _model_params(dataset, H, W, N):
  return types.SimpleNamespace(
      name_feature_extractor='resnet_v1_50', norm_layer='batch', norm_train_variables=True,
      batch_norm_accumulate_statistics=False, cross_replica_norm=False, psp_module=False, per_pixel_dataset_name=dataset,
      height_feature_extractor=H, width_feature_extractor=W, upsampling_method='bilinear', stride_feature_extractor=8,
      feature_dims_decreased=256, fov_expansion_kernel_rate=0, fov_expansion_kernel_size=0, Nb=N, distribute=False,
      regularization_weight=0.00017, batch_norm_decay=0.9, replace_voids=False, restore_emas=False, log_dir='/tmp/unused',
      init_ckpt_path='')


def main():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import tensorflow as tf
  from tensorflow import _slim, _train
  assert tf.__version__.endswith('shim')
  from estimator import define_estimator_hierarchical as de
  from models import resnet50_extended_model_hierarchical as rm
  config = types.SimpleNamespace(train_distribute=None, keep_checkpoint_max=2)
  out = {}

  def call(mode, features, labels, params):
    del _slim.REQUESTED[:], _slim.UPDATE_OPS[:], _slim.NORM_CALLS[:], _slim.REGULARIZED[:]
    _slim._unique.clear()
    tf.reset_collections()
    stdout, sys.stdout = sys.stdout, io.StringIO()
    try:
      with torch.no_grad():
        return de.define_estimator(mode, features, labels, rm.model, config, params)
    finally:
      sys.stdout = stdout

  for tag, (dataset, nbatches, N, H, W, LH, LW) in EVAL_CASES.items():
    _slim.reset(case_params(dataset))
    _train.reset()
    t2e = CITYSCAPES_TRAINING_CIDS2EVALUATION_CIDS if dataset == 'cityscapes' else VISTAS_TRAINING_CIDS2EVALUATION_CIDS
    params = _model_params(dataset, H, W, N)
    if tag.endswith('no_upsampling'):
      params.upsampling_method = 'no'
    params.training_cids2evaluation_cids = list(t2e)
    g = torch.Generator().manual_seed(SEED + len(tag))
    for b in range(nbatches):
      images = torch.rand(N, H, W, 3, generator=g) * 2 - 1
      blocks = torch.randint(0, len(t2e), (N, -(-LH // 4), -(-LW // 4)), generator=g, dtype=torch.int32)
      prolabels = blocks.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :LH, :LW].contiguous()
      spec = call(tf.estimator.ModeKeys.EVAL, {'proimages': tf.as_tf(images.clone())}, {'prolabels': prolabels}, params)
      assert spec.mode == tf.estimator.ModeKeys.EVAL and float(spec.loss) == 0.0
      out[f'{tag}/batch{b}/images'] = images.numpy()
      out[f'{tag}/batch{b}/prolabels'] = prolabels.numpy().astype(np.uint8)
      out[f'{tag}/batch{b}/decisions'] = torch.Tensor(spec.predictions['decisions']).numpy().astype(np.uint8)
      value, update_op = spec.eval_metric_ops['confusion_matrix']
      assert value.dtype == torch.int32
    out[f'{tag}/confusion_matrix'] = torch.Tensor(value).numpy().astype(np.int32)
    out[f'{tag}/num_classes'] = np.asarray(value.shape[0], dtype=np.int32)
    out[f'{tag}/training_cids2evaluation_cids'] = np.asarray(t2e, dtype=np.int32)
    out[f'{tag}/prediction_keys'] = np.asarray('\n'.join(sorted(spec.predictions.keys())))
    # evaluate_saver (define_savers.py:38-69): checkpoint name -> graph variable
    out[f'{tag}/saver'] = np.asarray('\n'.join(f'{k} {v.op.name}' for k, v in sorted(spec.scaffold.saver.var_list.items())))
    cm = out[f'{tag}/confusion_matrix']
    print(f'{tag}: cm sum {int(cm.sum())} trace {int(np.trace(cm))} classes {cm.shape[0]} keys {sorted(spec.predictions.keys())}')

  for tag, (dataset, N, H, W, system, raw) in PREDICT_CASES.items():
    _slim.reset(case_params(dataset))
    _train.reset()
    params = _model_params(dataset, H, W, N)
    params.restore_emas = raw is not None      # the second case restores the EMA shadows (predict_saver, --restore_emas)
    params.height_system, params.width_system = system
    g = torch.Generator().manual_seed(SEED + len(tag))
    images = torch.rand(N, H, W, 3, generator=g) * 2 - 1
    features = {'proimages': tf.as_tf(images.clone())}
    if raw is not None:
      features['rawimages'] = tf.as_tf(torch.randint(0, 256, (N, raw[0], raw[1], 3), generator=g, dtype=torch.int32).to(torch.uint8))
      features['rawimagespaths'] = ['a.png'] * N
    spec = call(tf.estimator.ModeKeys.PREDICT, features, None, params)
    out[f'{tag}/images'] = images.numpy()
    out[f'{tag}/size'] = np.asarray(tuple(spec.predictions['decisions'].shape[1:3]), dtype=np.int32)
    out[f'{tag}/prediction_keys'] = np.asarray('\n'.join(sorted(spec.predictions.keys())))
    out[f'{tag}/decisions'] = torch.Tensor(spec.predictions['decisions']).numpy().astype(np.uint8)
    out[f'{tag}/saver'] = np.asarray('\n'.join(f'{k} {v.op.name}' for k, v in sorted(spec.scaffold.saver.var_list.items())))
    for k in PROB_KEYS:
      out[f'{tag}/{k}'] = torch.Tensor(spec.predictions[k]).numpy().astype(np.float32)[:, ::PROB_STRIDE, ::PROB_STRIDE]
    print(f'{tag}: size {out[f"{tag}/size"].tolist()} keys {sorted(spec.predictions.keys())}')
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
  main()
