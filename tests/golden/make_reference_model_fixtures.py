"""Golden vectors of the NETWORK produced by running the reference's own `model()`:  tests/golden/reference_model_run.npz.

    python tests/golden/make_reference_model_fixtures.py       # needs /root/reference; run in the build container

`code/models/resnet50_extended_model_hierarchical.py::model` (with `feature_extractor`, `module_arg_scope`,
`_create_upsampler`, `_create_psp_module`) is imported UNMODIFIED with tests/golden/tf_shim first on sys.path and called
on seeded images; the variables it asks for come from a parameter dictionary keyed by TF variable names
(oracle.network.init_params: random numbers only).  tf.contrib.slim itself is third-party code the reference does not
carry: tests/golden/tf_shim/tensorflow/_slim.py restates the few slim / resnet_v1 functions that get called.  What
becomes a golden vector therefore is the reference's OWN wiring - which layers exist under which variable names, which
carry a normaliser / an activation, the arg-scope constants (epsilon 1e-5, decay, scale), the output-stride-8 ResNet
call, the extension / pyramid / field-of-view layers, the three adaptation bottlenecks, logits, upsampling, softmax /
argmax and the decision composition with the literal class-id tables of model():
  <tag>/images, <tag>/{l1,l2_vehicle,l2_human}_logits (every 3rd row / column), <tag>/{decisions,l1_decisions,l2_vehicle_decisions,l2_human_decisions}
  <tag>/variables (the names model() requested, in order), <tag>/regularized, <tag>/norm_calls, <tag>/params_checksum
and, for the training-mode case, the moving-statistic updates three layers would queue.
tests/test_reference_fixtures.py checks the oracle (and the product's variable names) against this file on CPU;
tests/test_gpu_reference_fixtures.py checks the CUDA network against it.
"""

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('WLSEG_REFERENCE', '/root/reference/code')
OUT = os.path.join(HERE, 'reference_model_run.npz')

# tag -> (dataset, N, H, W, mode is TRAIN, batch_norm_accumulate_statistics, init_params kwargs, model flags)
CASES = {
    'cs_eval': ('cityscapes', 1, 40, 56, False, False, {}, {}),
    'cs_train_bn': ('cityscapes', 2, 40, 56, True, True, {}, {}),
    'vistas_eval': ('vistas', 1, 40, 56, False, False, {}, {}),
    'cs_psp_fov_hybrid': ('cityscapes', 1, 48, 64, False, False, {'psp': True, 'fov': (3, 2), 'upsampling': 'hybrid'},
                          {'psp_module': True, 'fov_expansion_kernel_size': 3, 'fov_expansion_kernel_rate': 2,
                           'upsampling_method': 'hybrid'}),
    'cs_group': ('cityscapes', 2, 40, 56, False, False, {'norm': 'group'}, {'norm_layer': 'group'}),
    # sizes that are no multiple of 8 - the shape class of train.py's Vistas default 621 x 855 (621 = 5, 855 = 7 mod 8):
    # asymmetric SAME padding, pooling on odd maps, ceil-divided feature maps, x8 upsampling to an odd size
    'vistas_odd_size': ('vistas', 1, 45, 63, False, False, {}, {}),
    'cs_odd_size_train_bn': ('cityscapes', 2, 37, 51, True, True, {}, {}),
}
SEED = 11
LOGIT_STRIDE = 3


def case_params(tag):
  """The parameter dictionary of a case (the tests rebuild it with the same call)."""
  sys.path.insert(0, ROOT)
  from oracle import network as onet
  dataset, _, _, _, _, _, init_kw, _ = CASES[tag]
  return onet.init_params(dataset, seed=SEED, randomize_bn=True, tame=True, **init_kw)


def checksum(params):
  return float(sum(float(v.double().abs().sum()) for v in params.values()))


def case_images(tag):
  _, N, H, W, _, _, _, _ = CASES[tag]
  g = torch.Generator().manual_seed(SEED + len(tag))
  return torch.rand(N, H, W, 3, generator=g) * 2 - 1


def main():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import tensorflow as tf
  from tensorflow import _slim
  assert tf.__version__.endswith('shim')
  from models import resnet50_extended_model_hierarchical as rm
  out = {}
  for tag, (dataset, N, H, W, train, accumulate, _, flags) in CASES.items():
    tfp = case_params(tag)
    _slim.reset(tfp)
    params = types.SimpleNamespace(
        norm_layer='batch', norm_train_variables=True, batch_norm_accumulate_statistics=accumulate,
        regularization_weight=0.00017, batch_norm_decay=0.9, cross_replica_norm=False, psp_module=False,
        per_pixel_dataset_name=dataset, height_feature_extractor=H, width_feature_extractor=W, upsampling_method='bilinear',
        stride_feature_extractor=8, feature_dims_decreased=256, fov_expansion_kernel_rate=0, fov_expansion_kernel_size=0, Nb=N,
        distribute=False)
    for k, v in flags.items():
      setattr(params, k, v)
    config = types.SimpleNamespace(train_distribute=None)
    images = case_images(tag)
    mode = tf.estimator.ModeKeys.TRAIN if train else tf.estimator.ModeKeys.EVAL
    with torch.no_grad():
      _, _, pred = rm.model(mode, tf.as_tf(images.clone()), None, config, params)
    assert sorted(set(_slim.REQUESTED)) == sorted(tfp.keys()), (set(tfp) ^ set(_slim.REQUESTED))
    out[f'{tag}/images'] = images.numpy()
    # full-resolution logits on every third row / column (the file stays small; decisions are kept whole)
    for k in ('l1_logits', 'l2_vehicle_logits', 'l2_human_logits'):
      out[f'{tag}/{k}'] = torch.Tensor(pred[k]).numpy().astype(np.float32)[:, ::LOGIT_STRIDE, ::LOGIT_STRIDE]
    for k in ('decisions', 'l1_decisions', 'l2_vehicle_decisions', 'l2_human_decisions'):
      out[f'{tag}/{k}'] = torch.Tensor(pred[k].to(torch.int32)).numpy().astype(np.uint8)
    out[f'{tag}/variables'] = np.asarray('\n'.join(_slim.REQUESTED))
    out[f'{tag}/regularized'] = np.asarray('\n'.join(f'{n} {s:g}' for n, s in _slim.REGULARIZED))
    out[f'{tag}/norm_calls'] = np.asarray('\n'.join(' '.join(str(x) for x in c) for c in _slim.NORM_CALLS))
    out[f'{tag}/params_checksum'] = np.asarray(checksum(tfp))
    if train:
      # [TF-1.12] fused batch norm: moving <- moving - (1 - decay) * (moving - batch statistic), Bessel-corrected variance
      for scope, mean, var, decay in _slim.UPDATE_OPS[::20]:
        out[f'{tag}/update/{scope}/mean'] = mean.numpy()
        out[f'{tag}/update/{scope}/unbiased_variance'] = var.numpy()
        out[f'{tag}/update/{scope}/decay'] = np.asarray(decay)
    print(f'{tag}: {len(_slim.REQUESTED)} variables, {len(_slim.REGULARIZED)} regularised kernels, '
          f'logits max {float(torch.Tensor(pred["l1_logits"]).abs().max()):.3f}')
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
  main()
