"""Checkpoint naming / selection / warm-start matching rules (wlseg/checkpoints.py) against the reference's
savers and initialisers: code/estimator/define_savers.py:38-66, code/estimator/define_initializers.py:72-131,
code/estimator/define_estimator_hierarchical.py:96-111.  Name logic runs on CPU; the state round trip through
the parameter arenas needs the device (marked gpu)."""

import types

import numpy as np
import pytest
import torch


def _stub_params(psp=False):
  from wlseg import arch
  return types.SimpleNamespace(specs=arch.conv_specs((14, 7, 3), psp=psp))


def _imagenet_ckpt_vars():
  """Variable list of slim's resnet_v1_50 ImageNet checkpoint (names as tf.train.list_variables prints them)."""
  from wlseg import arch
  out = [('global_step', ()), ('resnet_v1_50/mean_rgb', (3,)), ('resnet_v1_50/logits/weights', (1, 1, 2048, 1000)),
         ('resnet_v1_50/logits/biases', (1000,))]
  for s in arch.conv_specs((14, 7, 3)):
    if '/resnet_v1_50/' not in s.scope:
      continue
    name = s.scope[len('feature_extractor/base/'):]
    out.append((f'{name}/weights', (s.R, s.S, s.C, s.K)))
    for v in ('beta', 'gamma', 'moving_mean', 'moving_variance'):
      out.append((f'{name}/BatchNorm/{v}', (s.K,)))
  return out


def test_model_variable_names_and_ema_selection():
  from wlseg import checkpoints as ck
  p = _stub_params()
  mv = ck.model_variables(p)
  assert len(mv) == 66 * 5
  names = [n for n, _ in mv]
  assert names[0] == 'feature_extractor/base/resnet_v1_50/conv1/weights' and mv[0][1] == (7, 7, 3, 64)
  assert 'feature_extractor/extension/decrease_fdims/BatchNorm/moving_variance' in names
  assert 'softmax_classifier/l2_human_logits/weights' in names
  # EMA shadows for everything but the BN moving statistics (define_estimator_hierarchical.py:103-106)
  assert sum(ck.has_ema(n) for n in names) == 66 * 3
  assert ck.ema_name('a/b/weights') == 'exponential_moving_averages/a/b/weights/ExponentialMovingAverage'


def test_predict_saver_keys():
  """define_savers.py:44-56: with --restore_emas every model variable except BatchNorm/moving_* is read from its
  EMA shadow; without it keys are the variable names."""
  from wlseg import checkpoints as ck
  p = _stub_params()
  plain = ck.predict_var_dict(p, restore_emas=False)
  assert all(k == v for k, v in plain.items()) and len(plain) == 330
  emas = ck.predict_var_dict(p, restore_emas=True)
  sc = 'adaptation_module/l1_features/conv2'
  assert emas[f'exponential_moving_averages/{sc}/weights/ExponentialMovingAverage'] == f'{sc}/weights'
  assert emas[f'exponential_moving_averages/{sc}/BatchNorm/gamma/ExponentialMovingAverage'] == f'{sc}/BatchNorm/gamma'
  assert emas[f'{sc}/BatchNorm/moving_mean'] == f'{sc}/BatchNorm/moving_mean'
  assert f'{sc}/weights' not in emas
  variables = {k: torch.zeros(1) for k in plain}   # a checkpoint trained without EMA
  with pytest.raises(KeyError, match='ExponentialMovingAverage'):
    ck.select_for_predict(p, variables, restore_emas=True)


@pytest.mark.parametrize('psp', [False, True])
def test_imagenet_warm_start_matching(psp):
  """define_initializers.py:92-115: exactly the 53 base-network convolutions (+ their BN variables) are
  initialised; extension / adaptation / classifier / pyramid layers, EMA shadows, Momentum slots and global_step
  never are, and the checkpoint's logits / mean_rgb find no home."""
  from wlseg import checkpoints as ck
  p = _stub_params(psp)
  mapping = ck.match_init_checkpoint(_imagenet_ckpt_vars(), ck.global_variables(p), psp_module=psp)
  assert len(mapping) == 53 * 5
  for cname, gname in mapping.items():
    assert gname == 'feature_extractor/base/' + cname
  assert not any(k.startswith(('resnet_v1_50/logits', 'resnet_v1_50/mean_rgb', 'global_step')) for k in mapping)
  # a checkpoint variable with the right name but another shape is not used (is_compatible_with)
  bad = [(n, (s[0], s[1], s[2], s[3] + 1) if len(s) == 4 else s) for n, s in _imagenet_ckpt_vars()]
  assert len(ck.match_init_checkpoint(bad, ck.global_variables(p), psp_module=psp)) == 53 * 4


def test_file_round_trip(tmp_path):
  from wlseg import checkpoints as ck
  v = {'a/weights': torch.randn(3, 3, 4, 8), 'a/BatchNorm/beta': torch.randn(8)}
  for ext in ('pt', 'npz'):
    path = ck.save_file(str(tmp_path / f'model.ckpt-7.{ext}'), v, 7)
    got, step = ck.load_file(path)
    assert step == 7 and set(got) == set(v)
    for k in v:
      assert torch.equal(got[k], v[k])
  # a bare {name: array} dict, the form the TF export script of INTEGRATION.md writes
  np.savez(str(tmp_path / 'imagenet.npz'), **{k: t.numpy() for k, t in v.items()})
  got, step = ck.load_file(str(tmp_path / 'imagenet.npz'))
  assert step == 0 and torch.equal(got['a/weights'], v['a/weights'])


@pytest.mark.gpu
def test_train_state_round_trip_and_restore_emas(cuda, tmp_path):
  """Two optimizer steps with EMA -> save -> (a) EVAL restore, plain and --restore_emas; (b) TRAIN resume with
  Momentum / EMA slots; (c) warm start of a fresh model from the saved base network."""
  import argparse
  from wlseg import checkpoints as ck, estimator as est, hierarchy, problem_defs
  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])

  def settings(**kw):
    s = argparse.Namespace(dtype='bf16', stride_feature_extractor=8, psp_module=False, batch_norm_decay=0.9,
                           momentum=0.9, use_nesterov=False, optimizer='SGDM', regularization_weight=0.00017,
                           ema_decay=0.9, distribute=False, learning_rate_schedule='piecewise_constant',
                           learning_rate_boundaries=[100], learning_rate_values=[0.01, 0.005], log_dir=str(tmp_path),
                           save_checkpoints_steps=None, restore_emas=False, init_ckpt_path=None)
    for k, v in kw.items():
      setattr(s, k, v)
    return s
  g = torch.Generator().manual_seed(1)

  def batches(n):
    for _ in range(n):
      yield ({'proimages': torch.rand(2, 64, 96, 3, generator=g) * 2 - 1},
             {'prolabels_per_pixel': torch.randint(0, 20, (2, 64, 96), generator=g, dtype=torch.int32)})
  e = est.Estimator(settings(), hier, device=cuda)
  e.initialize(log_dir=str(tmp_path), seed=3, for_training=True)
  e.train(batches(2), 2)
  path = e.save(str(tmp_path))
  variables, step = ck.load_file(path)
  assert step == 2 and len(variables) == 330 + 2 * 198
  w = 'feature_extractor/base/resnet_v1_50/block2/unit_1/bottleneck_v1/conv2/weights'
  assert tuple(variables[w].shape) == (3, 3, 128, 128)
  assert float((variables[ck.ema_name(w)] - variables[w]).abs().max()) > 0       # the shadow lags the variable
  assert float(variables[ck.momentum_name(w)].abs().max()) > 0
  # (a) EVAL restore
  plain = est.Estimator(settings(), hier, device=cuda)
  plain.initialize(ckpt_path=path)
  assert plain.global_step == 2 and torch.equal(plain.params.master, e.params.master)
  assert torch.equal(plain.params.moving, e.params.moving)
  emas = est.Estimator(settings(restore_emas=True), hier, device=cuda)
  emas.initialize(ckpt_path=path)
  assert torch.equal(emas.params.master, e.trainer.ws.ema_shadow) and torch.equal(emas.params.moving, e.params.moving)
  # (b) TRAIN resume: weights, bf16 operands, Momentum and EMA slots are exactly the saved ones, and the next
  # step gives the loss of the uninterrupted run (not bit-exact: atomics order, tests/test_gpu_train.py)
  r = est.Estimator(settings(), hier, device=cuda)
  r.initialize(log_dir=str(tmp_path), for_training=True)
  nxt = list(batches(1))
  r.train(iter(nxt), 0)   # creates the trainer and imports the slots, no step
  assert r.global_step == 2 and r.trainer.global_step == 2
  assert torch.equal(r.params.master, e.params.master) and torch.equal(r.params.operand, e.params.operand)
  assert torch.equal(r.trainer.ws.momentum, e.trainer.ws.momentum)
  assert torch.equal(r.trainer.ws.ema_shadow, e.trainer.ws.ema_shadow)
  lr_, le_ = r.train(iter(nxt), 1), e.train(iter(nxt), 1)
  assert r.global_step == 3 and np.isfinite(lr_).all()
  print('resumed vs uninterrupted losses', lr_[0].tolist(), le_[0].tolist())
  assert np.allclose(lr_[0], le_[0], rtol=5e-2)
  # (c) warm start from an "ImageNet" file holding the base network under slim's names (+ a foreign variable)
  base = {k[len('feature_extractor/base/'):]: v for k, v in variables.items() if k.startswith('feature_extractor/base/')}
  base['resnet_v1_50/logits/weights'] = torch.zeros(1, 1, 2048, 1000)
  init = ck.save_file(str(tmp_path / 'init' / 'resnet_v1_50.npz'), base, 0)
  fresh = est.Estimator(settings(init_ckpt_path=init, log_dir=str(tmp_path / 'empty')), hier, device=cuda)
  fresh.initialize(log_dir=str(tmp_path / 'empty'), seed=11, for_training=True)
  got = fresh.params.to_tf_dict()
  assert torch.equal(got[w], variables[w])
  d = 'feature_extractor/extension/decrease_fdims/weights'
  assert not torch.equal(got[d], variables[d])     # 'extension' is excluded: keeps its fresh random init
  assert float(got[d].abs().max()) > 0
