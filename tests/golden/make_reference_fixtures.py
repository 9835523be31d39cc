"""Golden vectors produced by RUNNING THE REFERENCE'S OWN PYTHON (not the oracle):  tests/golden/reference_run.npz.

    python tests/golden/make_reference_fixtures.py        # needs /root/reference; run in the build container

`/root/reference/code` is imported with tests/golden/tf_shim first on sys.path, so `import tensorflow` resolves to
the small eager emulation there (TensorFlow 1.12 itself cannot be installed: Python 3.12, no network).  The functions
below are the reference's own objects, called unmodified:

  estimator.define_losses_hierarchical.define_losses / _segment_sum      (define_losses_hierarchical.py:14-224)
  estimator.define_estimator_hierarchical._map_predictions_to_new_cids   (:490-528)
  estimator.define_estimator_hierarchical._resize_predictions            (:530-571)
  estimator.define_estimator_hierarchical._replace_voids                 (:573-630; on the key set it accepts)
  estimator.define_metrics.mean_iou                                      (define_metrics.py:5-20)
  estimator.define_optimizer.define_optimizer                            (define_optimizer.py:3-26)
  input_pipelines.open_images.input_subset_bboxes_v2._generate_rla       (input_subset_bboxes_v2.py:74-98)
  input_pipelines.open_images.input_subset_image_labels._generate_rla    (input_subset_image_labels.py:73-94)
  input_pipelines.utils.get_temp_Nb, from_0_1_to_m1_1                    (input_pipelines/utils.py:93-124)
  input_pipelines.utils.resize_images_and_labels                         (input_pipelines/utils.py:181-247)
  utils.utils._replacevoids, print_metrics_from_confusion_matrix         (utils/utils.py:286-289,385-446)

Inputs are seeded and stored next to the outputs; gradients of the reference's `total` loss with respect to the
LOW-RESOLUTION logits come from torch autograd through the shim (the bilinear x8 upsampling is the call
resnet50_extended_model_hierarchical.py:167 makes: tf.image.resize_images(..., align_corners=True)).
tests/test_reference_fixtures.py checks the oracle against this file on CPU; tests/test_gpu_reference_fixtures.py
checks the CUDA kernels against it on the GPU.  The file travels; /root/reference does not.
"""

import argparse
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('WLSEG_REFERENCE', '/root/reference/code')
OUT = os.path.join(HERE, 'reference_run.npz')


def import_reference():
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(0, REF)
  import tensorflow as tf
  assert tf.__version__.endswith('shim')
  from estimator import define_losses_hierarchical as dl
  from estimator import define_estimator_hierarchical as de
  from estimator import define_metrics as dm
  from estimator import define_optimizer as do
  from input_pipelines import utils as iu
  from input_pipelines.open_images import input_subset_bboxes_v2 as ib
  from utils import utils as uu
  return tf, dl, de, dm, do, iu, ib, uu


HEADS = {'cityscapes': (14, 7, 3, 20), 'vistas': (53, 12, 5, 66)}
MIDS = None


def make_boxes(rng, n_boxes, ib):
  """Open Images style annotation in the form `_imageid_and_bboxes_generator` yields (input_subset_bboxes_v2.py:
  56-72): byte-string mids + normalised (xmin, xmax, ymin, ymax).  One unknown mid is always included: the reference
  skips it."""
  mids_all = [m for m in ib.mid2cid.keys() if m != 'void']
  mids, coords = [], []
  for _ in range(n_boxes):
    mids.append(mids_all[rng.integers(0, len(mids_all))].encode('utf-8'))
    x = np.sort(rng.random(2))
    y = np.sort(rng.random(2))
    coords.append((float(x[0]), float(x[1]), float(y[0]), float(y[1])))
  mids.append('/m/unknown'.encode('utf-8'))
  coords.append((0.0, 1.0, 0.0, 1.0))
  return mids, np.asarray(coords, dtype=np.float32)


def losses_case(tf, dl, iu, ib, dataset, n_pp, n_pb, n_pi, h, w, seed, out, tag):
  c1, cv, ch, ncls = HEADS[dataset]
  H, W = 8 * h, 8 * w
  rng = np.random.default_rng(seed)
  g = torch.Generator().manual_seed(seed)
  nb = n_pp + n_pb + n_pi
  low = [(2.0 * torch.randn(nb, h, w, c, generator=g)).requires_grad_(True) for c in (c1, cv, ch)]
  # strong labels: 4x4 blocks of one class, voids included (the last id)
  blocks = torch.randint(0, ncls, (n_pp, H // 4, W // 4), generator=g, dtype=torch.int32)
  per_pixel = blocks.repeat_interleave(4, 1).repeat_interleave(4, 2).contiguous()
  # bbox labels through the reference's own rasteriser
  rla, box_lists = [], []
  for i in range(n_pb):
    mids, coords = make_boxes(rng, int(rng.integers(1, 7)), ib)
    rla.append(ib._generate_rla(b'img', mids, coords, np.asarray([H, W], dtype=np.int32)))
    cids = np.asarray([ib.mid2cid.get(m.decode('utf-8'), -1) for m in mids], dtype=np.int32)
    box_lists.append((coords, cids))
  per_bbox = torch.from_numpy(np.stack(rla).astype(np.float32)) if n_pb else torch.zeros(0, H, W, 15)
  # image-level labels as input_subset_image_labels.py:73-107 tiles them: m classes, value 1/m, spatially constant
  per_image = torch.zeros(n_pi, H, W, 15)
  for i in range(n_pi):
    m = int(rng.integers(1, 4))
    cls = rng.choice(14, size=m, replace=False)
    per_image[i, :, :, torch.as_tensor(cls)] = 1.0 / m
  # the model's upsampler and decision (resnet50_extended_model_hierarchical.py:84-93,167)
  full = [tf.image.resize_images(z, [H, W], align_corners=True) for z in low]
  l1_probs = tf.nn.softmax(full[0])
  l1_decs = tf.cast(tf.argmax(l1_probs, 3), tf.int32)
  predictions = {'l1_logits': full[0], 'l1_decisions': l1_decs,
                 'l2_vehicle_logits': full[1], 'l2_vehicle_probabilities': tf.nn.softmax(full[1]),
                 'l2_human_logits': full[2], 'l2_human_probabilities': tf.nn.softmax(full[2])}
  labels = {'prolabels_per_pixel': per_pixel, 'prolabels_per_bbox': per_bbox, 'prolabels_per_image': per_image}
  params = types.SimpleNamespace(Nb_per_pixel=n_pp, Nb_per_bbox=n_pb, Nb_per_image=n_pi, per_pixel_dataset_name=dataset)
  config = types.SimpleNamespace(train_distribute=None)
  tf.reset_collections()
  reg = torch.tensor(0.0625)
  tf.add_regularization_loss(reg)
  stdout, sys.stdout = sys.stdout, io.StringIO()   # the reference prints a notice about the 0.1 coefficient
  try:
    losses = dl.define_losses(tf.estimator.ModeKeys.TRAIN, predictions, labels, config, params)
  finally:
    sys.stdout = stdout
  losses['total'].backward()
  out[f'{tag}/dataset'] = np.asarray(dataset)
  out[f'{tag}/counts'] = np.asarray([n_pp, n_pb, n_pi, h, w], dtype=np.int32)
  for name, z in zip(('l1', 'l2_vehicle', 'l2_human'), low):
    out[f'{tag}/lowres_{name}_logits'] = z.detach().numpy()
    out[f'{tag}/grad_lowres_{name}_logits'] = z.grad.numpy()
  out[f'{tag}/l1_decisions'] = l1_decs.numpy()
  out[f'{tag}/prolabels_per_pixel'] = per_pixel.numpy()
  # the dense 15-channel weak labels are large: store float16 (exact: values are k/n with n <= 6 ... not exact in
  # general) -> keep fp32 but only for the small cases; boxes are stored as lists as well
  out[f'{tag}/prolabels_per_bbox'] = per_bbox.numpy()
  out[f'{tag}/prolabels_per_image_vectors'] = per_image[:, 0, 0, :].numpy()
  for i, (coords, cids) in enumerate(box_lists):
    out[f'{tag}/bbox{i}_coords'] = coords
    out[f'{tag}/bbox{i}_cids'] = cids
  for k in ('total', 'l1_segmentation', 'l1_segmentation_hot', 'l2_vehicle_segmentation', 'l2_human_segmentation',
            'regularization'):
    out[f'{tag}/loss_{k}'] = np.asarray(float(losses[k]), dtype=np.float32)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--out', default=OUT)
  args = ap.parse_args()
  tf, dl, de, dm, do, iu, ib, uu = import_reference()
  out = {}

  # ---- define_losses: mixed and strong-only, both label hierarchies
  losses_case(tf, dl, iu, ib, 'cityscapes', 2, 2, 1, 3, 4, 101, out, 'losses_cityscapes_mixed')
  losses_case(tf, dl, iu, ib, 'cityscapes', 2, 0, 0, 3, 5, 102, out, 'losses_cityscapes_strong')
  losses_case(tf, dl, iu, ib, 'vistas', 1, 2, 1, 4, 3, 103, out, 'losses_vistas_mixed')
  losses_case(tf, dl, iu, ib, 'vistas', 2, 0, 0, 2, 6, 104, out, 'losses_vistas_strong')

  # ---- _segment_sum on its own (the worked example of :112-113: one human + one vehicle box on a pixel)
  lab = torch.zeros(1, 1, 2, 15)
  lab[0, 0, 0, 2] = 0.5   # car
  lab[0, 0, 0, 6] = 0.5   # human
  lab[0, 0, 1, 14] = 1.0  # void
  veh_ids = tf.cast([5, 2, 0, 4, 3, 1, 6, 6, 6, 6, 6, 6, 6, 6, 6], tf.int32)
  out['segment_sum/labels'] = lab.numpy()
  out['segment_sum/ids'] = veh_ids.numpy()
  out['segment_sum/out'] = dl._segment_sum(lab, veh_ids, tf.reduce_max(veh_ids) + 1).numpy()

  # ---- _generate_rla: the normalisation examples of :87-95
  mids = [b'/m/0k4j', b'/m/0k4j', b'/m/01bjv', b'/m/01g317', b'/m/nothing']
  coords = np.asarray([(0.0, 0.5, 0.0, 0.5), (0.25, 0.75, 0.25, 0.75), (0.4, 0.9, 0.1, 0.6), (0.6, 1.0, 0.6, 1.0),
                       (0.0, 1.0, 0.0, 1.0)], dtype=np.float32)
  out['rla/coords'] = coords
  out['rla/cids'] = np.asarray([ib.mid2cid.get(m.decode(), -1) for m in mids], dtype=np.int32)
  out['rla/size'] = np.asarray([20, 28], dtype=np.int32)
  out['rla/out'] = ib._generate_rla(b'x', mids, coords, np.asarray([20, 28], dtype=np.int32))

  # ---- image-level labels: input_subset_image_labels._generate_rla (:73-94) - one, several, duplicated, unknown, no mids
  from input_pipelines.open_images import input_subset_image_labels as il
  cid2mid = {}
  for mid, cid in il.mid2cid.items():
    cid2mid.setdefault(cid, mid)
  cases = [[3], [0, 5, 9], [2, 2, 11], [], [13, 0]]
  vectors = []
  for k, cids in enumerate(cases):
    mids = [cid2mid[c].encode('utf-8') for c in cids] + ([b'/m/unknown'] if k == 2 else [])
    vectors.append(il._generate_rla(b'img', np.asarray(mids, dtype=object), np.asarray([4, 6], dtype=np.int32)))
  out['rla_image/cids'] = np.asarray([c + [-1] * (3 - len(c)) for c in cases], dtype=np.int32)
  out['rla_image/out'] = np.stack(vectors).astype(np.float32)

  # ---- _map_predictions_to_new_cids: the worked example (:494-496) and the Cityscapes evaluation map
  g = torch.Generator().manual_seed(7)
  probs5 = torch.softmax(torch.randn(1, 3, 4, 5, generator=g), -1)
  decs5 = torch.randint(0, 5, (1, 3, 4), generator=g, dtype=torch.int32)
  new = de._map_predictions_to_new_cids({'l1_probabilities': probs5, 'decisions': decs5}, [-1, 1, 1, 0, -1])
  out['remap/probs'] = probs5.numpy()
  out['remap/decisions'] = decs5.numpy()
  out['remap/map'] = np.asarray([-1, 1, 1, 0, -1], dtype=np.int32)
  out['remap/out_probs'] = new['l1_probabilities'].numpy()
  out['remap/out_decisions'] = new['decisions'].numpy()
  # hierarchical model: 14 l1 channels against a 20-entry map -> the probabilities are NOT transformed (:516-522)
  probs14 = torch.softmax(torch.randn(1, 2, 3, 14, generator=g), -1)
  decs20 = torch.randint(0, 20, (1, 2, 3), generator=g, dtype=torch.int32)
  cs_map = list(range(19)) + [-1]
  new = de._map_predictions_to_new_cids({'l1_probabilities': probs14, 'decisions': decs20}, cs_map)
  out['remap_cs/decisions'] = decs20.numpy()
  out['remap_cs/map'] = np.asarray(cs_map, dtype=np.int32)
  out['remap_cs/out_decisions'] = new['decisions'].numpy()
  out['remap_cs/probs_untouched'] = np.asarray(bool(torch.equal(new['l1_probabilities'], probs14)))

  # ---- _resize_predictions: up and down, odd sizes
  for tag, (ih, iw), (oh, ow) in (('resize_up', (13, 17), (31, 40)), ('resize_down', (24, 30), (11, 7))):
    p1 = torch.softmax(torch.randn(2, ih, iw, 14, generator=g), -1)
    pv = torch.softmax(torch.randn(2, ih, iw, 7, generator=g), -1)
    ph = torch.softmax(torch.randn(2, ih, iw, 3, generator=g), -1)
    d = torch.randint(0, 20, (2, ih, iw), generator=g, dtype=torch.int32)
    new = de._resize_predictions({'decisions': d.clone(), 'l1_probabilities': p1, 'l2_vehicle_probabilities': pv,
                                  'l2_human_probabilities': ph}, [oh, ow], None)
    out[f'{tag}/size'] = np.asarray([oh, ow], dtype=np.int32)
    for k, v in (('decisions', d), ('l1_probabilities', p1), ('l2_vehicle_probabilities', pv), ('l2_human_probabilities', ph)):
      out[f'{tag}/in_{k}'] = v.numpy()
      out[f'{tag}/out_{k}'] = new[k].numpy()

  # ---- _replace_voids on a flat classifier (the only key set it accepts, :589-592)
  p = torch.softmax(2 * torch.randn(1, 6, 7, 5, generator=g), -1)
  d = tf.cast(tf.argmax(p, 3), tf.int32)
  new = de._replace_voids({'l1_probabilities': p, 'decisions': d}, None)
  out['replace_voids/probs'] = p.numpy()
  out['replace_voids/decisions'] = d.numpy()
  out['replace_voids/out_decisions'] = new['decisions'].to(torch.int32).numpy()

  # ---- define_metrics.mean_iou
  lab = torch.randint(0, 20, (2, 16, 24), generator=g, dtype=torch.int32)
  dec = torch.where(torch.rand(2, 16, 24, generator=g) < 0.6, lab, torch.randint(0, 20, (2, 16, 24), generator=g, dtype=torch.int32))
  out['mean_iou/labels'] = lab.numpy()
  out['mean_iou/decisions'] = dec.numpy()
  out['mean_iou/out'] = np.asarray(float(dm.mean_iou(lab, dec, 20, None)), dtype=np.float32)

  # ---- define_optimizer: the Cityscapes default schedule of system_factory.py:207-233 and the polynomial one
  prm = types.SimpleNamespace(learning_rate_schedule='piecewise_constant', learning_rate_boundaries=[8 * 743, 15 * 743],
                              learning_rate_values=[0.01, 0.005, 0.0025], optimizer='SGDM', momentum=0.9, use_nesterov=False)
  steps = [0, 1, 8 * 743 - 1, 8 * 743, 8 * 743 + 1, 15 * 743, 15 * 743 + 1, 17 * 743]
  out['lr/steps'] = np.asarray(steps, dtype=np.int64)
  out['lr/piecewise'] = np.asarray([do.define_optimizer(s, prm).learning_rate for s in steps], dtype=np.float64)
  prm2 = types.SimpleNamespace(learning_rate_schedule='polynomial_decay', learning_rate_initial=0.01, num_training_steps=17 * 743,
                               learning_rate_final=0.0001, learning_rate_power=0.9, optimizer='SGDM', momentum=0.9, use_nesterov=True)
  out['lr/polynomial'] = np.asarray([do.define_optimizer(s, prm2).learning_rate for s in steps], dtype=np.float64)
  # three momentum updates through the optimizer object define_optimizer returns (plain and Nesterov)
  w0 = torch.randn(16, generator=g)
  grads = [torch.randn(16, generator=g) for _ in range(3)]
  out['sgdm/w0'] = w0.numpy()
  out['sgdm/grads'] = torch.stack(grads).numpy()
  for name, p_ in (('plain', prm), ('nesterov', types.SimpleNamespace(**{**vars(prm), 'use_nesterov': True}))):
    opt = do.define_optimizer(0, p_)
    w = w0.clone()
    for gr in grads:
      w = opt.apply_dense('w', w, gr)
    out[f'sgdm/{name}'] = w.numpy()

  # ---- small host helpers
  out['replacevoids/in'] = np.asarray([-1, 1, 1, 0, -1], dtype=np.int32)
  out['replacevoids/out'] = np.asarray(uu._replacevoids([-1, 1, 1, 0, -1]), dtype=np.int32)
  out['temp_nb/out'] = np.asarray([iu.get_temp_Nb(types.SimpleNamespace(train_distribute=None), 8),
                                   iu.get_temp_Nb(types.SimpleNamespace(train_distribute=types.SimpleNamespace(num_towers=4)), 8)],
                                  dtype=np.int32)
  out['m1_1/out'] = iu.from_0_1_to_m1_1(torch.tensor([0.0, 0.25, 0.5, 1.0])).numpy()
  # print_metrics_from_confusion_matrix: a cm with an empty row (class excluded from the means) and a union-0 class
  cm = np.asarray([[5, 1, 0, 0], [2, 7, 0, 1], [0, 0, 0, 0], [1, 0, 0, 3]], dtype=np.int32)
  buf = io.StringIO()
  with np.errstate(all='ignore'):
    uu.print_metrics_from_confusion_matrix(cm, printfile=buf, summary=True)
  out['print_metrics/cm'] = cm
  out['print_metrics/summary'] = np.asarray(buf.getvalue())

  # ---- resize_images_and_labels (input_pipelines/utils.py:181-247): aspect-preserving resize + random crop
  for tag, shape, lab_kind, target, preserve, seed in (
      ('crop_ids', (2, 20, 31, 3), 'ids', (16, 16), True, 5),
      ('crop_dense', (1, 24, 18, 3), 'dense', (20, 20), True, 9),
      ('crop_ids_tall', (1, 37, 16, 3), 'ids', (12, 12), True, 2),
      ('resize_plain', (2, 20, 31, 3), 'ids', (13, 40), False, 1)):
    gg = torch.Generator().manual_seed(seed)
    img = torch.rand(*shape, generator=gg)
    if lab_kind == 'ids':
      lab = torch.randint(0, 20, shape[:3], generator=gg, dtype=torch.int32)
    else:
      lab = torch.softmax(3 * torch.randn(*shape[:3], 15, generator=gg), -1)
    tf.set_random_seed(seed)
    pi, pl = iu.resize_images_and_labels(tf.as_tf(img), tf.as_tf(lab), target, preserve_aspect_ratio=preserve)
    out[f'{tag}/images'] = img.numpy()
    out[f'{tag}/labels'] = lab.numpy()
    out[f'{tag}/target'] = np.asarray(target, dtype=np.int32)
    out[f'{tag}/preserve'] = np.asarray(preserve)
    out[f'{tag}/offset'] = np.asarray(list(tf.RANDOM_LOG) if preserve else [0, 0], dtype=np.int32)
    out[f'{tag}/out_images'] = torch.Tensor(pi).numpy() if False else pi.as_subclass(torch.Tensor).numpy()
    out[f'{tag}/out_labels'] = pl.as_subclass(torch.Tensor).numpy()

  np.savez_compressed(args.out, **out)
  print(f'wrote {args.out}: {len(out)} arrays, {os.path.getsize(args.out) / 1024:.0f} KB')


if __name__ == '__main__':
  main()
