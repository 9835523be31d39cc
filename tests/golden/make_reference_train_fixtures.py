"""Golden vectors of the TRAINING STEP produced by running the reference's own `define_estimator` (TRAIN branch):
tests/golden/reference_train_run.npz.

    python tests/golden/make_reference_train_fixtures.py       # needs /root/reference; run in the build container

`code/estimator/define_estimator_hierarchical.py::define_estimator` is imported UNMODIFIED with tests/golden/tf_shim
first on sys.path and called in TRAIN mode with the reference's own `model()` as `model_fn`, once per optimizer step
(the shim is eager: one call = building the graph and running `train_op` once; variables, the global step, Momentum
slots and EMA shadows live in the shim's stores across calls, as they live in the TF session).  What runs is the
reference's code for (define_estimator_hierarchical.py:77-159):
  model() in training mode (batch statistics, moving-statistic updates into UPDATE_OPS, l2 regularisers),
  define_losses under `losses/` (strong + bbox + image-level batch),
  the ExponentialMovingAverage block (which variables, decay, num_updates=global_step, UPDATE_OPS),
  define_optimizer (piecewise-constant schedule crossing a boundary inside the run, Momentum),
  create_train_op(losses['total'], optimizer, global_step) and the EstimatorSpec.
Restated (TF itself is un-vendored): tensorflow/_train.py - create_train_op's ordering (UPDATE_OPS before the
gradient application, global step last), ExponentialMovingAverage's formula.  One reference function is replaced by a
no-op for the run: `_define_summaries` (TensorBoard images / scalars, out of scope in DESIGN.md section 0).

Stored per case <tag>:
  images, prolabels_per_pixel, bbox lists (coords, cids), image-level vectors, the schedule;
  step<i>/losses (total, segmentation?, l1, l2_vehicle, l2_human, regularization), step<i>/learning_rate, step<i>/global_step
  final/<variable> for a handful of variables (weights, gamma / beta, moving statistics) + their Momentum slots and
  EMA shadows, and final/checksums: sum |x| over every variable / slot / shadow (float64), in `names` order;
  final/update_checksums: sum |final - initial| per variable.
tests/test_reference_fixtures.py replays the run with the oracle on CPU; tests/test_gpu_reference_fixtures.py with the
product's Trainer (fp32 check mode and the bf16 product path) on the GPU.
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
OUT = os.path.join(HERE, 'reference_train_run.npz')

SEED = 23
# Conditioning: a train-mode batch-norm ResNet at plain random init is a strongly amplifying map (two fp32 evaluations
# of the SAME graph differ by 2e-2 in the gradient, tests/test_gpu_train.py; after two updates the oracle in fp32 and in
# fp64 disagree by 3-18 % in the losses).  `case_params` therefore scales gamma of every residual branch's last
# normaliser by 0.1 (a near zero-init-residual start): the oracle's own fp32-vs-fp64 deviation after three steps at the
# reference's default learning rate 0.01 drops to 1e-4, and a trajectory comparison means something.  The second case
# uses a large regularisation weight so that the l2 term is visible in the updates next to the data gradient.
RESIDUAL_GAMMA_SCALE = 0.1
CASES = {
    'cs_mixed_sgdm_ema': ('cityscapes', 1, 1, 1, 40, 56, 3,
                          dict(learning_rate_schedule='piecewise_constant', learning_rate_boundaries=[1],
                               learning_rate_values=[0.01, 0.005], optimizer='SGDM', momentum=0.9, use_nesterov=False,
                               ema_decay=0.9, regularization_weight=0.00017, batch_norm_decay=0.9)),
    'cs_strong_nesterov_poly': ('cityscapes', 2, 0, 0, 32, 48, 2,
                                dict(learning_rate_schedule='polynomial_decay', learning_rate_initial=0.02,
                                     learning_rate_final=0.0001, learning_rate_power=0.9, num_training_steps=5,
                                     optimizer='SGDM', momentum=0.8, use_nesterov=True, ema_decay=0.0,
                                     regularization_weight=0.02, batch_norm_decay=0.95)),
    'vistas_mixed_sgdm': ('vistas', 1, 1, 1, 40, 56, 2,
                          dict(learning_rate_schedule='piecewise_constant', learning_rate_boundaries=[50],
                               learning_rate_values=[0.01, 0.005], optimizer='SGDM', momentum=0.9, use_nesterov=False,
                               ema_decay=0.0, regularization_weight=0.00017, batch_norm_decay=0.9)),
    'cs_psp_fov_hybrid': ('cityscapes', 2, 0, 0, 48, 64, 2,
                          dict(learning_rate_schedule='piecewise_constant', learning_rate_boundaries=[50],
                               learning_rate_values=[0.01, 0.005], optimizer='SGDM', momentum=0.9, use_nesterov=False,
                               ema_decay=0.9, regularization_weight=0.00017, batch_norm_decay=0.9,
                               model=({'psp': True, 'fov': (3, 2), 'upsampling': 'hybrid'},
                                      {'psp_module': True, 'fov_expansion_kernel_size': 3, 'fov_expansion_kernel_rate': 2,
                                       'upsampling_method': 'hybrid'}))),
    # sizes that are no multiple of 8 (the shape class of train.py's Vistas default 621 x 855): asymmetric SAME padding,
    # the strided unit and both pooling layers on odd maps, and their gradients
    'cs_odd_size_momentum': ('cityscapes', 1, 1, 0, 37, 51, 2,
                             dict(learning_rate_schedule='piecewise_constant', learning_rate_boundaries=[50],
                                  learning_rate_values=[0.01, 0.005], optimizer='SGDM', momentum=0.9, use_nesterov=False,
                                  ema_decay=0.0, regularization_weight=0.00017, batch_norm_decay=0.9)),
    'cs_group_norm': ('cityscapes', 1, 1, 0, 40, 56, 2,
                      dict(learning_rate_schedule='piecewise_constant', learning_rate_boundaries=[50],
                           learning_rate_values=[0.01, 0.005], optimizer='SGD', momentum=0.9, use_nesterov=False,
                           ema_decay=0.0, regularization_weight=0.00017, batch_norm_decay=0.9,
                           model=({'norm': 'group'}, {'norm_layer': 'group'}))),
}
# variables stored whole (everything else: checksums)
KEEP = ('feature_extractor/base/resnet_v1_50/conv1/weights',
        'feature_extractor/base/resnet_v1_50/conv1/BatchNorm/gamma',
        'feature_extractor/base/resnet_v1_50/conv1/BatchNorm/moving_mean',
        'feature_extractor/base/resnet_v1_50/conv1/BatchNorm/moving_variance',
        'feature_extractor/base/resnet_v1_50/block3/unit_2/bottleneck_v1/conv2/BatchNorm/beta',
        'feature_extractor/base/resnet_v1_50/block4/unit_3/bottleneck_v1/conv3/BatchNorm/moving_variance',
        'adaptation_module/l2_vehicle_features/conv1/weights',
        'softmax_classifier/l1_logits/weights',
        'softmax_classifier/l2_human_logits/weights',
        'softmax_classifier/l2_human_logits/BatchNorm/beta')


def case_params(tag):
  """The initial parameter dictionary of a case (the tests rebuild it with the same call)."""
  if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
  from oracle import network as onet
  init_kw = CASES[tag][7].get('model', ({}, {}))[0]
  params = onet.init_params(CASES[tag][0], seed=SEED, randomize_bn=True, tame=True, **init_kw)
  for k in params:
    if k.endswith(('/conv3/BatchNorm/gamma', '/conv3/GroupNorm/gamma')):
      params[k] = params[k] * RESIDUAL_GAMMA_SCALE
  return params


def case_batches(tag, ib=None):
  """Seeded images and labels of every step.  Box lists are (coords[n, 4] normalised xmin xmax ymin ymax, cids[n])."""
  dataset, n_pp, n_pb, n_pi, H, W, steps, _ = CASES[tag]
  ncls = 20 if dataset == 'cityscapes' else 66
  g = torch.Generator().manual_seed(SEED + 3 * len(tag))
  rng = np.random.default_rng(SEED + len(tag))
  batches = []
  for _ in range(steps):
    images = torch.rand(n_pp + n_pb + n_pi, H, W, 3, generator=g) * 2 - 1
    blocks = torch.randint(0, ncls, (n_pp, -(-H // 8), -(-W // 8)), generator=g, dtype=torch.int32)
    per_pixel = blocks.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :H, :W].contiguous()
    boxes = []
    for _ in range(n_pb):
      n = int(rng.integers(2, 6))
      cids = rng.integers(0, 14, size=n).astype(np.int32)
      x = np.sort(rng.random((n, 2)), axis=1)
      y = np.sort(rng.random((n, 2)), axis=1)
      boxes.append((np.concatenate([x, y], axis=1).astype(np.float32), cids))
    vectors = np.zeros((n_pi, 15), dtype=np.float32)
    for i in range(n_pi):
      m = int(rng.integers(1, 4))
      vectors[i, rng.choice(14, size=m, replace=False)] = 1.0 / m
    batches.append((images, per_pixel, boxes, vectors))
  return batches


def imagenet_checkpoint_variables(model_names, shapes):
  """What tf.train.list_variables reports for the slim ImageNet resnet_v1_50 checkpoint --init_ckpt_path points at: the
  base network's variables without the reference's `feature_extractor/base/` prefix, plus the checkpoint's own
  global_step, mean_rgb and 1000-way logits."""
  out = [('global_step', ()), ('resnet_v1_50/mean_rgb', (3,)), ('resnet_v1_50/logits/weights', (1, 1, 2048, 1000)),
         ('resnet_v1_50/logits/biases', (1000,))]
  prefix = 'feature_extractor/base/'
  for n in model_names:
    if n.startswith(prefix + 'resnet_v1_50/'):
      out.append((n[len(prefix):], tuple(shapes[n])))
  return sorted(out)


def warm_start_cases(tf, _slim, _train, de, rm, out):
  """The TRAIN graph built with --init_ckpt_path set (nothing is run): `replace_initializers`
  (define_initializers.py:72-131) chooses which graph variables the ImageNet checkpoint initialises, `train_saver`
  (define_savers.py:3-36) which variables are saved; stored next to tf.global_variables() of that graph."""
  if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
  from oracle import network as onet
  ckpt = '/imagenet/resnet_v1_50.ckpt'
  for tag, psp in (('warm_start', False), ('warm_start_psp', True)):
    tfp = onet.init_params('cityscapes', seed=SEED, randomize_bn=True, tame=True, psp=psp)
    for n in tfp:
      if '/moving_' not in n:
        tfp[n].requires_grad_(True)
    _slim.reset(tfp)
    _train.reset()
    tf.reset_collections()
    _train.CHECKPOINT_VARIABLES[ckpt] = imagenet_checkpoint_variables(sorted(tfp), {k: v.shape for k, v in tfp.items()})
    H, W = 48, 64
    params = types.SimpleNamespace(
        name_feature_extractor='resnet_v1_50', norm_layer='batch', norm_train_variables=True,
        batch_norm_accumulate_statistics=True, cross_replica_norm=False, psp_module=psp, per_pixel_dataset_name='cityscapes',
        height_feature_extractor=H, width_feature_extractor=W, upsampling_method='bilinear', stride_feature_extractor=8,
        feature_dims_decreased=256, fov_expansion_kernel_rate=0, fov_expansion_kernel_size=0, Nb=1, Nb_per_pixel=1,
        Nb_per_bbox=0, Nb_per_image=0, distribute=False, init_ckpt_path=ckpt, log_dir='/tmp/unused', num_training_steps=100,
        save_checkpoints_steps=50, learning_rate_schedule='piecewise_constant', learning_rate_boundaries=[50],
        learning_rate_values=[0.01, 0.005], optimizer='SGDM', momentum=0.9, use_nesterov=False, ema_decay=0.9,
        regularization_weight=0.00017, batch_norm_decay=0.9)
    config = types.SimpleNamespace(train_distribute=None, keep_checkpoint_max=2)
    g = torch.Generator().manual_seed(SEED)
    features = {'proimages': tf.as_tf(torch.rand(1, H, W, 3, generator=g) * 2 - 1)}
    labels = {'prolabels_per_pixel': torch.randint(0, 20, (1, H, W), generator=g, dtype=torch.int32),
              'prolabels_per_bbox': torch.zeros(0, H, W, 15), 'prolabels_per_image': torch.zeros(0, H, W, 15)}
    stdout, sys.stdout = sys.stdout, io.StringIO()
    try:
      spec = de.define_estimator(tf.estimator.ModeKeys.TRAIN, features, labels, rm.model, config, params)
    finally:
      sys.stdout = stdout
    (path, var_dict), = _train.INIT_FROM_CHECKPOINT
    assert path == ckpt
    out[f'{tag}/checkpoint_variables'] = np.asarray('\n'.join(f'{n} {" ".join(str(d) for d in s)}' for n, s in _train.CHECKPOINT_VARIABLES[ckpt]))
    out[f'{tag}/global_variables'] = np.asarray('\n'.join(v.op.name for v in tf.global_variables()))
    out[f'{tag}/init_from_checkpoint'] = np.asarray('\n'.join(f'{k} {v.op.name}' for k, v in sorted(var_dict.items())))
    out[f'{tag}/train_saver'] = np.asarray('\n'.join(v.op.name for v in spec.scaffold.saver.var_list))
    print(f'{tag}: {len(tf.global_variables())} global variables, {len(var_dict)} initialised from the checkpoint, '
          f'{len(spec.scaffold.saver.var_list)} saved')


def main():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import tensorflow as tf
  from tensorflow import _slim, _train
  assert tf.__version__.endswith('shim')
  from estimator import define_estimator_hierarchical as de
  from input_pipelines.open_images import input_subset_bboxes_v2 as ib
  from models import resnet50_extended_model_hierarchical as rm
  de._define_summaries = lambda *a, **k: None     # TensorBoard summaries: out of scope (module docstring)
  captured = {}
  reference_define_losses = de.define_losses

  def recording_define_losses(*a, **k):              # the reference's function, its returned dictionary kept for the file
    captured.clear()
    captured.update(reference_define_losses(*a, **k))
    return captured
  de.define_losses = recording_define_losses
  cid2mid = {}
  for mid, cid in ib.mid2cid.items():
    cid2mid.setdefault(cid, mid)
  out = {}
  for tag, (dataset, n_pp, n_pb, n_pi, H, W, steps, opt) in CASES.items():
    tfp = case_params(tag)
    names = sorted(tfp.keys())
    trainable = [n for n in names if '/moving_' not in n]
    for n in trainable:
      tfp[n].requires_grad_(True)
    _slim.reset(tfp)
    _train.reset()
    params = types.SimpleNamespace(
        name_feature_extractor='resnet_v1_50', norm_layer='batch', norm_train_variables=True,
        batch_norm_accumulate_statistics=True, cross_replica_norm=False, psp_module=False, per_pixel_dataset_name=dataset,
        height_feature_extractor=H, width_feature_extractor=W, upsampling_method='bilinear', stride_feature_extractor=8,
        feature_dims_decreased=256, fov_expansion_kernel_rate=0, fov_expansion_kernel_size=0, Nb=n_pp + n_pb + n_pi,
        Nb_per_pixel=n_pp, Nb_per_bbox=n_pb, Nb_per_image=n_pi, distribute=False, init_ckpt_path='', log_dir='/tmp/unused',
        num_training_steps=opt.get('num_training_steps', 100), save_checkpoints_steps=50,
        **{k: v for k, v in opt.items() if k not in ('num_training_steps', 'model')})
    for k, v in opt.get('model', ({}, {}))[1].items():
      setattr(params, k, v)
    config = types.SimpleNamespace(train_distribute=None, keep_checkpoint_max=2)
    batches = case_batches(tag)
    out[f'{tag}/names'] = np.asarray('\n'.join(names))
    for i, (images, per_pixel, boxes, vectors) in enumerate(batches):
      # weak labels in the dense form the input pipeline hands to the model function: the reference's own rasteriser
      rla = []
      for coords, cids in boxes:
        mids = [cid2mid[int(c)].encode('utf-8') for c in cids]
        rla.append(ib._generate_rla(b'img', mids, coords, np.asarray([H, W], dtype=np.int32)))
      per_bbox = torch.from_numpy(np.stack(rla).astype(np.float32)) if n_pb else torch.zeros(0, H, W, 15)
      per_image = torch.from_numpy(vectors)[:, None, None, :].expand(n_pi, H, W, 15).contiguous() if n_pi else torch.zeros(0, H, W, 15)
      features = {'proimages': tf.as_tf(images.clone())}
      labels = {'prolabels_per_pixel': per_pixel, 'prolabels_per_bbox': per_bbox, 'prolabels_per_image': per_image}
      # one session.run(train_op): per-call graph state is rebuilt, the stores persist
      del _slim.REQUESTED[:], _slim.UPDATE_OPS[:], _slim.NORM_CALLS[:], _slim.REGULARIZED[:]
      _slim._unique.clear()
      _train.COLLECTIONS.clear()
      tf.reset_collections()
      step_before = int(tf.train.get_or_create_global_step())
      stdout, sys.stdout = sys.stdout, io.StringIO()
      try:
        spec = de.define_estimator(tf.estimator.ModeKeys.TRAIN, features, labels, rm.model, config, params)
      finally:
        printed, sys.stdout = sys.stdout.getvalue(), stdout
      assert spec.mode == tf.estimator.ModeKeys.TRAIN and spec.scaffold.saver is None
      lr = float(spec.train_op.optimizer.learning_rate)
      total = float(spec.train_op())
      assert abs(total - float(spec.loss)) == 0.0
      ld = captured
      out[f'{tag}/step{i}/images'] = images.numpy()
      out[f'{tag}/step{i}/prolabels_per_pixel'] = per_pixel.numpy().astype(np.uint8)
      for j, (coords, cids) in enumerate(boxes):
        out[f'{tag}/step{i}/bbox{j}_coords'] = coords
        out[f'{tag}/step{i}/bbox{j}_cids'] = cids
      out[f'{tag}/step{i}/image_vectors'] = vectors
      out[f'{tag}/step{i}/learning_rate'] = np.asarray(lr, dtype=np.float64)
      out[f'{tag}/step{i}/global_step_before'] = np.asarray(step_before, dtype=np.int64)
      out[f'{tag}/step{i}/losses'] = np.asarray([total] + [float(ld[k]) for k in (
          'l1_segmentation', 'l2_vehicle_segmentation', 'l2_human_segmentation', 'regularization')], dtype=np.float64)
      print(f'{tag} step {i}: global_step {step_before} lr {lr:g} losses {out[f"{tag}/step{i}/losses"]}')
      if i == 0:
        found = [ln for ln in printed.splitlines() if ln.startswith('Found ')]
        out[f'{tag}/ema_notice'] = np.asarray(found[0] if found else '')
    assert int(tf.train.get_or_create_global_step()) == steps
    out[f'{tag}/global_step'] = np.asarray(steps, dtype=np.int64)
    sums = lambda d, key: np.asarray([float(d[key(n)].detach().double().abs().sum()) if key(n) in d else -1.0  # noqa: E731
                                      for n in names], dtype=np.float64)
    out[f'{tag}/final/checksums'] = sums(_slim.VARS, lambda n: n)
    initial = case_params(tag)
    out[f'{tag}/final/update_checksums'] = np.asarray([float((_slim.VARS[n].detach() - initial[n]).double().abs().sum())
                                                       for n in names], dtype=np.float64)
    out[f'{tag}/final/momentum_checksums'] = sums(_train.OPT_SLOTS, lambda n: n)
    out[f'{tag}/final/ema_checksums'] = sums(_train.EMA_SHADOWS, lambda n: f'exponential_moving_averages/{n}/ExponentialMovingAverage')
    out[f'{tag}/ema_names'] = np.asarray('\n'.join(sorted(_train.EMA_SHADOWS.keys())))
    for n in KEEP:
      if n not in _slim.VARS:        # group norm: no moving statistics, <scope>/GroupNorm/{beta,gamma}
        n = n.replace('/BatchNorm/', '/GroupNorm/')
        if n not in _slim.VARS:
          continue
      out[f'{tag}/final/{n}'] = _slim.VARS[n].detach().numpy().copy()
      if n in _train.OPT_SLOTS:
        out[f'{tag}/final_momentum/{n}'] = _train.OPT_SLOTS[n].numpy().copy()
      sh = f'exponential_moving_averages/{n}/ExponentialMovingAverage'
      if sh in _train.EMA_SHADOWS:
        out[f'{tag}/final_ema/{n}'] = _train.EMA_SHADOWS[sh].numpy().copy()
    print(f'{tag}: {len(names)} variables, {len(_train.OPT_SLOTS)} Momentum slots, {len(_train.EMA_SHADOWS)} EMA shadows')
  warm_start_cases(tf, _slim, _train, de, rm, out)
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
  main()
