"""Oracle forward pass: ResNet-50 (OS8) + extension + adaptation + hierarchical heads.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned against the reference's own
`model()` run over tests/golden/tf_shim (tests/golden/reference_model_run.npz: identical
decisions, logits to 0 .. 1e-4); slim's `resnet_v1_50` internals themselves are restated.

Follows, in order:
  code/models/resnet50_extended_feature_extractor.py:8-51   (base + decrease_fdims)
  code/models/resnet50_extended_model_hierarchical.py:17-141 (adaptation, logits,
      upsample, softmax, argmax, hierarchical composition)
  [TF-1.12] slim resnet_v1 / resnet_utils semantics, SURVEY.md section 3.4 + Appendix A.

Parameters live in a flat dict keyed by the TF variable names
(`.../weights` HWIO, `.../BatchNorm/{gamma,beta,moving_mean,moving_variance}`).
"""

import collections

import torch

from oracle import tfops
from oracle.tables import TABLES

# (block name, base depth, number of units, nominal stride of the block)
_BLOCKS = [('block1', 64, 3, 2), ('block2', 128, 4, 2), ('block3', 256, 6, 2), ('block4', 512, 3, 1)]
_RES = 'feature_extractor/base/resnet_v1_50'


PSP_SCOPES = tuple('feature_extractor/pyramid_module/Conv' + ('' if i == 0 else f'_{i}') for i in range(5))
PSP_BINS = (1, 2, 3, 6)


FOV_SCOPE = 'feature_extractor/extension/increase_fov'
# --upsampling_method hybrid: one slim.conv2d_transpose per head inside variable_scope('upsampling'), default
# scopes Conv2d_transpose, Conv2d_transpose_1, _2 (resnet50_extended_model_hierarchical.py:84-86,161-179)
UPSAMPLING_SCOPES = tuple('softmax_classifier/upsampling/Conv2d_transpose' + ('' if i == 0 else f'_{i}') for i in range(3))


def conv_specs(dataset='cityscapes', feature_dims_decreased=256, psp=False, fov=None):
  """Ordered {scope: (kh, kw, cin, cout)} for the 66 convs of the default model (+5 with --psp_module:
  slim's default scopes Conv, Conv_1 .. Conv_4 under feature_extractor/pyramid_module,
  resnet50_extended_model_hierarchical.py:55-57,186-207)."""
  c1, cv, ch = TABLES[dataset]['head_widths']
  specs = collections.OrderedDict()
  specs[f'{_RES}/conv1'] = (7, 7, 3, 64)
  cin = 64
  for name, base, units, _ in _BLOCKS:
    for u in range(1, units + 1):
      sc = f'{_RES}/{name}/unit_{u}/bottleneck_v1'
      if cin != base * 4:
        specs[f'{sc}/shortcut'] = (1, 1, cin, base * 4)
      specs[f'{sc}/conv1'] = (1, 1, cin, base)
      specs[f'{sc}/conv2'] = (3, 3, base, base)
      specs[f'{sc}/conv3'] = (1, 1, base, base * 4)
      cin = base * 4
  d = feature_dims_decreased
  specs['feature_extractor/extension/decrease_fdims'] = (1, 1, cin, d)
  if fov:  # --fov_expansion_kernel_size / _rate (resnet50_extended_feature_extractor.py:44-49)
    specs[FOV_SCOPE] = (fov[0], fov[0], d, d)
  if psp:
    for sc in PSP_SCOPES[:4]:
      specs[sc] = (1, 1, d, d)
    specs[PSP_SCOPES[4]] = (1, 1, 5 * d, d)
  for br in ('l1_features', 'l2_vehicle_features', 'l2_human_features'):
    sc = f'adaptation_module/{br}'
    specs[f'{sc}/conv1'] = (1, 1, d, d)
    specs[f'{sc}/conv2'] = (3, 3, d, d)
    specs[f'{sc}/conv3'] = (1, 1, d, d)
  specs['softmax_classifier/l1_logits'] = (1, 1, d, c1)
  specs['softmax_classifier/l2_vehicle_logits'] = (1, 1, d, cv)
  specs['softmax_classifier/l2_human_logits'] = (1, 1, d, ch)
  return specs


def init_params(dataset='cityscapes', seed=0, randomize_bn=False, tame=False, psp=False, fov=None, upsampling='bilinear',
                norm='batch'):
  """Random init as the reference's arg scope does (variance scaling conv
  kernels, gamma=1, beta=0, moving_mean=0, moving_var=1;
  resnet50_extended_model_hierarchical.py:335-340).  `randomize_bn=True`
  perturbs the BN variables so that tests exercise the BN arithmetic; `tame=True`
  halves gamma on every residual-branch output / shortcut BN so that activations
  of the random network stay O(1) through the 16 units (inference-mode BN with
  moving stats 0/1 is otherwise an identity and the residual sums grow)."""
  g = torch.Generator().manual_seed(seed)
  params = collections.OrderedDict()
  for scope, shape in conv_specs(dataset, psp=psp, fov=fov).items():
    params[f'{scope}/weights'] = tfops.variance_scaling_trunc_normal(shape, g)
    c = shape[3]
    if randomize_bn:
      params[f'{scope}/BatchNorm/gamma'] = 0.75 + 0.5 * torch.rand(c, generator=g)
      params[f'{scope}/BatchNorm/beta'] = 0.1 * torch.randn(c, generator=g)
      params[f'{scope}/BatchNorm/moving_mean'] = 0.1 * torch.randn(c, generator=g)
      params[f'{scope}/BatchNorm/moving_variance'] = 0.75 + 0.5 * torch.rand(c, generator=g)
    else:
      params[f'{scope}/BatchNorm/gamma'] = torch.ones(c)
      params[f'{scope}/BatchNorm/beta'] = torch.zeros(c)
      params[f'{scope}/BatchNorm/moving_mean'] = torch.zeros(c)
      params[f'{scope}/BatchNorm/moving_variance'] = torch.ones(c)
    if tame and (scope.endswith('/conv3') or scope.endswith('/shortcut')):
      params[f'{scope}/BatchNorm/gamma'] = params[f'{scope}/BatchNorm/gamma'] * 0.5
    if norm == 'group':
      # --norm_layer group: tf.contrib.layers.group_norm variables <scope>/GroupNorm/{beta,gamma}, no moving statistics
      for v in ('gamma', 'beta'):
        params[f'{scope}/GroupNorm/{v}'] = params.pop(f'{scope}/BatchNorm/{v}')
      del params[f'{scope}/BatchNorm/moving_mean'], params[f'{scope}/BatchNorm/moving_variance']
  if upsampling == 'hybrid':
    for sc, c in zip(UPSAMPLING_SCOPES, TABLES[dataset]['head_widths']):
      params[f'{sc}/weights'] = tfops.variance_scaling_trunc_normal((3, 3, c, c), g)   # [kh, kw, out, in]
      params[f'{sc}/biases'] = (0.1 * torch.randn(c, generator=g)) if randomize_bn else torch.zeros(c)
  return params


class _RoundSte(torch.autograd.Function):
  """bf16 storage rounding with a straight-through gradient."""

  @staticmethod
  def forward(ctx, x):
    return x.to(torch.bfloat16).to(x.dtype)

  @staticmethod
  def backward(ctx, g):
    return g


class _RoundGrad(torch.autograd.Function):
  """Identity whose incoming gradient is rounded to bf16 (a gradient tensor stored in bf16)."""

  @staticmethod
  def forward(ctx, x):
    return x.view_as(x)

  @staticmethod
  def backward(ctx, g):
    return g.to(torch.bfloat16).to(g.dtype)


class Net:
  """Functional forward over a parameter dict.

  storage='bf16' restates the SAME graph with the storage roundings of the product path made
  explicit (images, conv kernels, pre-BN conv outputs and activations rounded to bf16 with
  straight-through gradients; batch statistics taken from the bf16-stored conv output; activation and
  pre-BN gradients rounded to bf16; the logits layers stay fp32).  A train-mode batch-norm network
  at random init amplifies a 1e-3 perturbation by ~10^2 (measured: tests/test_gpu_train.py), so
  comparing a bf16 pipeline with the fp32 graph end to end says nothing; comparing it with the
  same graph rounded at the same points does.

  training=True uses batch statistics in every BN layer (the reference's
  `batch_norm_accumulate_statistics`, train.py:45-46) and records the updated
  moving statistics in `self.new_moving`.
  """

  def __init__(self, params, dataset='cityscapes', training=False, bn_decay=0.9, eps=1e-5, storage='fp32',
               psp=False, fov=None, upsampling='bilinear', norm='batch'):
    assert storage in ('fp32', 'bf16') and upsampling in ('bilinear', 'hybrid', 'no') and norm in ('batch', 'group')
    assert norm == 'batch' or storage == 'fp32', 'the bf16-storage restatement covers batch norm only'
    self.norm = norm
    self.upsampling = upsampling
    self.psp = psp
    self.fov = fov  # (kernel size, dilation rate) of extension/increase_fov, or None
    self.storage = storage
    self.p = params
    self.dataset = dataset
    self.training = training
    self.bn_decay = bn_decay
    self.eps = eps
    self.new_moving = {}
    self.taps = {}
    self.record_layers = False
    self.layer_taps = {}  # scope -> (pre-BN conv output, layer output), filled when record_layers

  def _bn(self, x, scope):
    if self.norm == 'group':
      # module_arg_scope: groups = 32, and 1 under softmax_classifier (resnet50_extended_model_hierarchical.py:75-77);
      # the same computation in every mode
      groups = 1 if scope.startswith('softmax_classifier/') else 32
      return tfops.group_norm(x, self.p[f'{scope}/GroupNorm/gamma'], self.p[f'{scope}/GroupNorm/beta'], groups, self.eps)
    bn = f'{scope}/BatchNorm'
    y, mm, mv, _, _ = tfops.batch_norm(
        x, self.p[f'{bn}/gamma'], self.p[f'{bn}/beta'],
        self.p[f'{bn}/moving_mean'], self.p[f'{bn}/moving_variance'],
        self.training, self.bn_decay, self.eps)
    if self.training:
      self.new_moving[f'{bn}/moving_mean'] = mm
      self.new_moving[f'{bn}/moving_variance'] = mv
    return y

  def _q(self, x):
    return _RoundSte.apply(x) if self.storage == 'bf16' else x

  def _qg(self, x):
    return _RoundGrad.apply(x) if self.storage == 'bf16' else x

  def _conv_bn(self, x, scope, stride=1, rate=1, relu=True, same_explicit=False, residual=None, fp32_out=False):
    """conv -> batch norm (-> + residual) (-> ReLU); the residual add sits here so that the bf16
    storage mode rounds the activation once, after the add, as the fused kernels do."""
    w = self._q(self.p[f'{scope}/weights'])
    if same_explicit:
      z = tfops.conv2d_same(x, w, stride, rate)
    else:
      z = tfops.conv2d(x, w, stride, rate, 'SAME')
    if self.storage == 'bf16' and self.training:
      z = self._qg(z)
      bn = f'{scope}/BatchNorm'
      n = z.shape[0] * z.shape[1] * z.shape[2]
      zs = z if fp32_out else self._q(z)
      # the product takes the batch statistics of the STORED (bf16) conv output - the tensor the
      # normalisation is applied to (csrc/conv_igemm_sm100.cu); gradients pass straight through the rounding
      mean = zs.mean(dim=(0, 1, 2))
      var = ((zs - mean) ** 2).mean(dim=(0, 1, 2))
      y = (zs - mean) * torch.rsqrt(var + self.eps) * self.p[f'{bn}/gamma'] + self.p[f'{bn}/beta']
      with torch.no_grad():
        self.new_moving[f'{bn}/moving_mean'] = self.p[f'{bn}/moving_mean'] - (1.0 - self.bn_decay) * (
            self.p[f'{bn}/moving_mean'] - mean)
        self.new_moving[f'{bn}/moving_variance'] = self.p[f'{bn}/moving_variance'] - (1.0 - self.bn_decay) * (
            self.p[f'{bn}/moving_variance'] - var * (n / max(n - 1, 1)))
    else:
      y = self._bn(z, scope)
    if residual is not None:
      y = y + residual
    if relu:
      y = torch.relu(y)
    if not fp32_out:
      y = self._qg(self._q(y))
    if self.record_layers:
      self.layer_taps[scope] = (z.detach(), y.detach())
    return y

  def _bottleneck(self, x, scope, depth, depth_bottleneck, stride, rate):
    """slim resnet_v1.bottleneck [TF-1.12]."""
    if x.shape[-1] == depth:
      shortcut = x if stride == 1 else tfops.max_pool_same(x, 1, stride)
    else:
      shortcut = self._conv_bn(x, f'{scope}/shortcut', stride=stride, relu=False)
    r = self._conv_bn(x, f'{scope}/conv1')
    r = self._conv_bn(r, f'{scope}/conv2', stride=stride, rate=rate, same_explicit=True)
    return self._conv_bn(r, f'{scope}/conv3', relu=True, residual=shortcut)

  def features(self, images, output_stride=8):
    """feature_extractor(): base resnet_v1_50(global_pool=False, output_stride)
    then extension/decrease_fdims.  images: NHWC fp32 in [-1, 1)."""
    x = self._conv_bn(self._q(images), f'{_RES}/conv1', stride=2, same_explicit=True)
    x = tfops.max_pool_same(x, 3, 2)
    self.taps['pool1'] = x
    # stack_blocks_dense: the root counts as stride 4
    current_stride, rate = 4, 1
    for name, base, units, block_stride in _BLOCKS:
      for u in range(1, units + 1):
        unit_stride = block_stride if u == units else 1
        sc = f'{_RES}/{name}/unit_{u}/bottleneck_v1'
        if current_stride == output_stride:
          x = self._bottleneck(x, sc, base * 4, base, 1, rate)
          rate *= unit_stride
        else:
          x = self._bottleneck(x, sc, base * 4, base, unit_stride, 1)
          current_stride *= unit_stride
      self.taps[name] = x
    x = self._conv_bn(x, 'feature_extractor/extension/decrease_fdims')
    self.taps['decrease_fdims'] = x
    if self.fov:
      # slim.conv2d(fe, C, kernel_size, rate=rate): stride 1, SAME padding, BN + ReLU from the arg scope
      x = self._conv_bn(x, FOV_SCOPE, rate=self.fov[1])
      self.taps['increase_fov'] = x
    if self.psp:
      x = self._psp(x)
      self.taps['pyramid_module'] = x
    return x

  def _psp(self, bottom):
    """_create_psp_module (resnet50_extended_model_hierarchical.py:186-207): VALID average pooling into
    bins {1, 2, 3, 6} (kernel = stride = feature size // bins), 1x1 conv (+BN+ReLU from the arg scope),
    bilinear align_corners back to the feature size, concat with the input, 1x1 conv."""
    h, w = bottom.shape[1], bottom.shape[2]
    outs = [bottom]
    for sc, b in zip(PSP_SCOPES[:4], PSP_BINS):
      k = (h // b, w // b)
      pooled = self._qg(self._q(tfops.avg_pool_valid(bottom, k, k)))
      c = self._conv_bn(pooled, sc)
      outs.append(self._qg(self._q(tfops.resize_bilinear(c, h, w, align_corners=True))))
    return self._conv_bn(torch.cat(outs, -1), PSP_SCOPES[4])

  def lowres_logits(self, images):
    f = self.features(images)
    d = f.shape[-1]
    out = []
    for br, lg in (('l1_features', 'l1_logits'), ('l2_vehicle_features', 'l2_vehicle_logits'),
                   ('l2_human_features', 'l2_human_logits')):
      a = self._bottleneck(f, f'adaptation_module/{br}', d, d, 1, 1)
      # slim.conv2d(activation_fn=None) inside the arg scope: BN still applied
      out.append(self._conv_bn(a, f'softmax_classifier/{lg}', relu=False, fp32_out=True))
    if self.upsampling == 'hybrid':
      # _create_upsampler 'hybrid': 3x3 transposed convolution (stride 1, + bias, fp32) before the resize
      out = [tfops.conv2d_transpose_same(z, self.p[f'{sc}/weights'], self.p[f'{sc}/biases'])
             for z, sc in zip(out, UPSAMPLING_SCOPES)]
    return out

  def forward(self, images):
    """model(): returns the 10-key predictions dict of
    resnet50_extended_model_hierarchical.py:121-130 (+ 'lowres_logits')."""
    hf, wf = images.shape[1], images.shape[2]
    low = self.lowres_logits(images)
    t = TABLES[self.dataset]
    if self.upsampling == 'no':   # `upsampled = bottom`: everything downstream stays at the feature resolution
      l1, l2v, l2h = low
    else:
      l1, l2v, l2h = [tfops.resize_bilinear(z, hf, wf, align_corners=True) for z in low]
    pred = compose_predictions(l1, l2v, l2h, self.dataset)
    pred['lowres_logits'] = low
    del t
    return pred


def compose_predictions(l1_logits, l2v_logits, l2h_logits, dataset):
  """softmax x3, argmax-of-probabilities x3, hierarchical composition
  (resnet50_extended_model_hierarchical.py:88-117)."""
  t = TABLES[dataset]
  l1_probs = tfops.softmax(l1_logits)
  l2v_probs = tfops.softmax(l2v_logits)
  l2h_probs = tfops.softmax(l2h_logits)
  l1_decs = tfops.argmax_first(l1_probs)
  l2v_decs = tfops.argmax_first(l2v_probs)
  l2h_decs = tfops.argmax_first(l2h_probs)
  l1_map = torch.tensor(t['l1_cids2common_cids'], dtype=torch.int32)
  v_map = torch.tensor(t['l2_vehicle_cids2common_cids'], dtype=torch.int32)
  h_map = torch.tensor(t['l2_human_cids2common_cids'], dtype=torch.int32)
  decs = torch.where(
      l1_decs == t['cid_l1_vehicle'], v_map[l2v_decs.long()],
      torch.where(l1_decs == t['cid_l1_human'], h_map[l2h_decs.long()], l1_map[l1_decs.long()]))
  return {'l1_logits': l1_logits, 'l1_probabilities': l1_probs, 'l1_decisions': l1_decs,
          'l2_vehicle_logits': l2v_logits, 'l2_vehicle_probabilities': l2v_probs,
          'l2_vehicle_decisions': l2v_decs,
          'l2_human_logits': l2h_logits, 'l2_human_probabilities': l2h_probs,
          'l2_human_decisions': l2h_decs,
          'decisions': decs}
