"""On-device synthetic input generator of Cityscapes / Vistas / Open Images *shapes*.

Replaces code/input_pipelines/** (tf.data over TFRecords / JPEGs / pickles that are not shipped)
and honours its tensor contract:
  features['proimages']            (Npp+Npb+Npi, hf, wf, 3) fp32 in [-1, 1), order strong, bbox, image
  labels['prolabels_per_pixel']    (Npp, hf, wf) int32 in [0, Ncls-1], void = last id
  labels['prolabels_per_bbox']     (Npb, hf, wf, 15) fp32 per-pixel multinomial (overlapping boxes
                                   normalised, void = 1 where no box), as `_generate_rla` of
                                   code/input_pipelines/open_images/input_subset_bboxes_v2.py:74-98
  labels['prolabels_per_image']    (Npi, hf, wf, 15) fp32 spatially constant, uniform over the present
                                   classes (input_subset_image_labels.py:73-107)
  eval labels['prolabels']         (N, hf, wf) int32
Seeded with 1234 + rank.  This is input plumbing in torch, not the hot path.
"""

import torch

NUM_WEAK = 15


class SyntheticInputs:
  def __init__(self, num_classes, device, rank=0, seed=1234):
    self.num_classes = num_classes
    self.device = torch.device(device)
    self.gen = torch.Generator(device=self.device)
    self.gen.manual_seed(seed + rank)

  def images(self, n, h, w):
    return torch.rand((n, h, w, 3), generator=self.gen, device=self.device, dtype=torch.float32) * 2.0 - 1.0

  def strong_labels(self, n, h, w, block=32, void_fraction=0.1):
    bh, bw = -(-h // block), -(-w // block)
    cls = torch.randint(0, self.num_classes, (n, bh, bw), generator=self.gen, device=self.device, dtype=torch.int32)
    void = torch.rand((n, bh, bw), generator=self.gen, device=self.device) < void_fraction
    cls = torch.where(void, torch.full_like(cls, self.num_classes - 1), cls)
    full = cls.repeat_interleave(block, 1).repeat_interleave(block, 2)
    return full[:, :h, :w].contiguous()

  def bbox_lists(self, n, max_boxes=12):
    """Compact Open Images-style annotation: per image k ~ U{1..max_boxes} boxes of class ~ U{0..13} with
    uniform corners; -> (coords fp32 [n, max_boxes, 4] = xmin, xmax, ymin, ymax normalised,
    cids int32 [n, max_boxes], -1 = padding) - the form input_subset_bboxes_v2.py:60-72 yields."""
    k = torch.randint(1, max_boxes + 1, (n, 1), generator=self.gen, device=self.device)
    cids = torch.randint(0, NUM_WEAK - 1, (n, max_boxes), generator=self.gen, device=self.device, dtype=torch.int32)
    slot = torch.arange(max_boxes, device=self.device).unsqueeze(0)
    cids = torch.where(slot < k, cids, torch.full_like(cids, -1))
    c = torch.rand((n, max_boxes, 4), generator=self.gen, device=self.device)
    xs, _ = torch.sort(c[..., 0:2], dim=-1)
    ys, _ = torch.sort(c[..., 2:4], dim=-1)
    return torch.cat([xs, ys], dim=-1).contiguous(), cids.contiguous()

  def bbox_labels(self, n, h, w, max_boxes=12):
    """Dense per-pixel multinomials rasterised ON THE DEVICE from the box lists
    (csrc/weak_labels.cu; bit-exact with `_generate_rla`, input_subset_bboxes_v2.py:74-98)."""
    from wlseg import ops
    coords, cids = self.bbox_lists(n, max_boxes)
    return ops.rasterize_bbox_labels(coords, cids, h, w)

  def image_vectors(self, n, max_classes=3):
    """Compact image-level labels: m ~ U{1..max_classes} classes per image, value 1/m -> fp32 [n, 15]."""
    vec = torch.zeros((n, NUM_WEAK), dtype=torch.float32, device=self.device)
    for i in range(n):
      m = int(torch.randint(1, max_classes + 1, (1,), generator=self.gen, device=self.device))
      cids = torch.randperm(NUM_WEAK - 1, generator=self.gen, device=self.device)[:m]
      vec[i, cids] = 1.0 / m
    return vec

  def image_labels(self, n, h, w, max_classes=3):
    """m ~ U{1..max_classes} classes per image, value 1/m, tiled over the image on the device
    (input_subset_image_labels.py:73-107)."""
    from wlseg import ops
    return ops.tile_image_labels(self.image_vectors(n, max_classes), h, w)

  # ---- batches in the reference's (features, labels) form ------------------------------------
  def train_batch(self, npp, npb, npi, h, w, compact=False):
    """compact: the weak labels stay in the form the Open Images side stores - (class, box) lists and one 15-way
    vector per image-level image ('bbox_coords', 'bbox_cids', 'image_vectors') - and the loss kernel expands them
    per pixel in registers (wlseg_loss_fwd_bwd_lists) instead of reading 60 B/pixel of dense labels."""
    features = {'proimages': self.images(npp + npb + npi, h, w)}
    labels = {'prolabels_per_pixel': self.strong_labels(npp, h, w)}
    if compact:
      if npb:
        labels['bbox_coords'], labels['bbox_cids'] = self.bbox_lists(npb)
      if npi:
        labels['image_vectors'] = self.image_vectors(npi)
    else:
      labels['prolabels_per_bbox'] = self.bbox_labels(npb, h, w) if npb else None
      labels['prolabels_per_image'] = self.image_labels(npi, h, w) if npi else None
    return features, labels

  def eval_batch(self, n, h, w):
    return {'proimages': self.images(n, h, w)}, {'prolabels': self.strong_labels(n, h, w)}


def eval_input_fn(config, params):
  """Drop-in `input_fn(config, params)` for SemanticSegmentation({'eval': ...}): yields
  `num_eval_steps` batches of Nb synthetic images; rank `r` of `R` takes every R-th batch."""
  del config
  rank, world = getattr(params, 'rank', 0), getattr(params, 'world_size', 1)
  src = SyntheticInputs(params.output_Nclasses, getattr(params, 'device', 'cuda'), rank=rank)
  for step in range(params.num_eval_steps):
    if step % world != rank:
      continue
    yield src.eval_batch(params.Nb, params.height_feature_extractor, params.width_feature_extractor)


def train_input_fn(config, params):
  """Synthetic counterpart of heterogeneous_supervision/per_pixel_per_bbox_per_image.train_input."""
  del config
  rank = getattr(params, 'rank', 0)
  src = SyntheticInputs(params.output_Nclasses, getattr(params, 'device', 'cuda'), rank=rank)
  from wlseg.estimator import get_temp_Nb
  npp = get_temp_Nb(params, params.Nb_per_pixel)
  npb = get_temp_Nb(params, params.Nb_per_bbox)
  npi = get_temp_Nb(params, params.Nb_per_image)
  while True:
    yield src.train_batch(npp, npb, npi, params.height_feature_extractor, params.width_feature_extractor)


def predict_input_fn(config, params):
  """Synthetic counterpart of dataset_agnostic_predict_input.predict_input: one image per batch,
  with `rawimages` uint8 and `rawimagespaths`."""
  del config
  src = SyntheticInputs(params.output_Nclasses, getattr(params, 'device', 'cuda'))
  n = getattr(params, 'steps', None) or 4
  h, w = params.height_feature_extractor, params.width_feature_extractor
  for i in range(n):
    pro = src.images(params.Nb, h, w)
    raw = ((pro + 1.0) * 127.5).clamp(0, 255).to(torch.uint8)
    yield {'proimages': pro, 'rawimages': raw, 'rawimagespaths': [f'synthetic_{i:05d}_{j}.png'.encode()
                                                                  for j in range(params.Nb)]}, None
