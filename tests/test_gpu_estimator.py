"""Estimator-level behaviour the reference gets from tf.estimator (code/system_factory.py:256-412,
code/estimator/define_estimator_hierarchical.py:96-129), through the public facade on the device:

* train() always leaves a checkpoint of its last step behind (CheckpointSaverHook.end) and evaluate() restores it;
* EVAL / PREDICT refuse to run on random-init weights unless --synthetic asks for them;
* the EMA shadows are updated BEFORE the gradient step with num_updates = the pre-increment global_step.
"""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W = 64, 96


def _train_settings(tmp_path, extra=()):
  from wlseg import settings as wsettings
  argv = [str(tmp_path), 'cityscapes', '--height_feature_extractor', str(H), '--width_feature_extractor', str(W),
          '--Nb', '2', '--steps', '3', '--synthetic'] + list(extra)
  st = wsettings.build_parser(wsettings.TRAIN).parse_args(argv)
  wsettings.train_extra_args(st)   # the reference's hard overrides (512 x 1024, 4 + 8 + 4 images) ...
  st.height_feature_extractor, st.width_feature_extractor = H, W   # ... replaced by a test-sized problem
  st.Nb_per_pixel, st.Nb_per_bbox, st.Nb_per_image, st.Nb = 2, 1, 1, 2
  st.device, st.rank, st.world_size = 'cuda:0', 0, 1
  return st


def _eval_settings(tmp_path, synthetic=True, extra=()):
  from wlseg import problem_defs, settings as wsettings
  argv = [str(tmp_path), '4', problem_defs.default_path('cityscapes'), 'synthetic', 'cityscapes', '--Nb', '2',
          '--height_feature_extractor', str(H), '--width_feature_extractor', str(W)] + (['--synthetic'] if synthetic else []) + list(extra)
  st = wsettings.eval_extra_args(wsettings.build_parser(wsettings.EVAL).parse_args(argv))
  st.device, st.rank, st.world_size = 'cuda:0', 0, 1
  return st


def test_train_leaves_final_checkpoint_and_evaluate_restores_it(cuda, tmp_path):
  from wlseg import synthetic
  from wlseg.system_factory import SemanticSegmentation
  # default cadence = one epoch (743 steps): no periodic save can fire within 3 steps
  system = SemanticSegmentation({'train': synthetic.train_input_fn}, None, _train_settings(tmp_path))
  losses = system.train()
  assert losses.shape == (3, 6) and np.isfinite(losses).all()
  ckpt = os.path.join(str(tmp_path), 'model.ckpt-3.pt')
  assert os.path.isfile(ckpt), sorted(os.listdir(str(tmp_path)))
  trained = system.estimator.params.master.clone()
  moving = system.estimator.params.moving.clone()
  # evaluate.py WITHOUT --synthetic: must find and restore that checkpoint (not random weights)
  ev = SemanticSegmentation({'eval': synthetic.eval_input_fn}, None, _eval_settings(tmp_path, synthetic=False))
  metrics = ev.evaluate()
  assert metrics[0]['global_step'] == 3
  assert torch.equal(ev.estimator.params.master, trained) and torch.equal(ev.estimator.params.moving, moving)
  assert int(metrics[0]['confusion_matrix'].sum()) <= 4 * H * W


def test_periodic_save_is_not_duplicated(cuda, tmp_path):
  from wlseg import synthetic
  from wlseg.system_factory import SemanticSegmentation
  system = SemanticSegmentation({'train': synthetic.train_input_fn}, None,
                                _train_settings(tmp_path, ['--save_checkpoints_steps', '3']))
  system.train()
  assert sorted(f for f in os.listdir(str(tmp_path)) if f.startswith('model.ckpt')) == ['model.ckpt-3.pt']


def test_eval_and_predict_without_checkpoint_raise(cuda, tmp_path):
  from wlseg import synthetic
  from wlseg.system_factory import SemanticSegmentation
  ev = SemanticSegmentation({'eval': synthetic.eval_input_fn}, None, _eval_settings(tmp_path, synthetic=False))
  with pytest.raises(ValueError, match='Could not find trained model'):
    ev.evaluate()
  # --synthetic opts into random-init weights (benchmarks)
  ev = SemanticSegmentation({'eval': synthetic.eval_input_fn}, None, _eval_settings(tmp_path, synthetic=True))
  assert int(ev.evaluate()[0]['confusion_matrix_int64'].sum()) == 4 * H * W


def test_ema_runs_before_the_gradient_step_with_preincrement_num_updates(cuda):
  """Known answer: shadow_0 = w_0.  Step 0 averages w_0 into itself (no change, d_0 = min(0.9, 1/10));
  step 1 averages w_1 (the weights after ONE update) with d_1 = min(0.9, 2/11): shadow = w_0 - (1 - d_1)(w_0 - w_1)."""
  from wlseg import hierarchy, network, problem_defs, trainer as wtrainer

  class S:
    momentum, use_nesterov, optimizer, regularization_weight = 0.9, False, 'SGDM', 0.00017
    batch_norm_decay, distribute, ema_decay = 0.9, False, 0.9

  hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
  params = network.Params(hier, cuda)
  params.init_random(1)
  w0 = params.master.clone()
  tr = wtrainer.Trainer(params, S, use_graph=False)
  g = torch.Generator().manual_seed(2)
  img = (torch.rand(2, H, W, 3, generator=g) * 2 - 1).to(cuda)
  lab = {'prolabels_per_pixel': torch.randint(0, 20, (2, H, W), generator=g, dtype=torch.int32).to(cuda)}
  tr.step({'proimages': img}, lab, 0.01)
  assert torch.equal(tr.ws.ema_shadow, w0)
  w1 = params.master.clone()
  assert not torch.equal(w1, w0)
  tr.step({'proimages': img}, lab, 0.01)
  d1 = min(0.9, 2.0 / 11.0)
  want = w0 - (1.0 - d1) * (w0 - w1)
  assert torch.allclose(tr.ws.ema_shadow, want, rtol=1e-6, atol=1e-7)
  # the returned loss vector is the caller's own copy (graph mode would otherwise overwrite it)
  tr2 = wtrainer.Trainer(params, S, use_graph=True)
  outs = [tr2.step({'proimages': img}, lab, 0.01) for _ in range(5)]
  torch.cuda.synchronize()
  assert len({o.data_ptr() for o in outs}) == 5
