"""Layer plan of the network the reference builds (names follow the TF variable scopes).

Restates, as data, the graph of
  code/models/resnet50_extended_feature_extractor.py:8-51  (slim resnet_v1_50, global_pool=False,
      output_stride=8, + extension/decrease_fdims)
  code/models/resnet50_extended_model_hierarchical.py:59-86 (three adaptation bottlenecks and the
      three logits convolutions, every convolution followed by batch norm)
with the [TF-1.12] slim semantics of SURVEY.md section 3.4: the block stride sits on the LAST unit's
3x3 conv; once the output stride is reached strides turn into dilation rates (block3: 2,
block4: 4); `conv2d_same` pads explicitly for stride 2; pool1 uses SAME padding.
"""

import collections

RES = 'feature_extractor/base/resnet_v1_50'
BLOCKS = (('block1', 64, 3, 2), ('block2', 128, 4, 2), ('block3', 256, 6, 2), ('block4', 512, 3, 1))
BRANCHES = (('l1_features', 'l1_logits'), ('l2_vehicle_features', 'l2_vehicle_logits'),
            ('l2_human_features', 'l2_human_logits'))

ConvSpec = collections.namedtuple('ConvSpec', 'scope R S C K stride dilation relu')
UnitSpec = collections.namedtuple('UnitSpec', 'scope depth_in depth bottleneck stride rate has_shortcut_conv')


def units(output_stride=8):
  """The 16 bottleneck units of resnet_v1_50 after stack_blocks_dense(output_stride)."""
  out = []
  current_stride, rate = 4, 1  # the root (conv1 + pool1) has stride 4
  cin = 64
  for name, base, n_units, block_stride in BLOCKS:
    for u in range(1, n_units + 1):
      unit_stride = block_stride if u == n_units else 1
      scope = f'{RES}/{name}/unit_{u}/bottleneck_v1'
      if current_stride == output_stride:
        out.append(UnitSpec(scope, cin, base * 4, base, 1, rate, cin != base * 4))
        rate *= unit_stride
      else:
        out.append(UnitSpec(scope, cin, base * 4, base, unit_stride, 1, cin != base * 4))
        current_stride *= unit_stride
      cin = base * 4
  return out


def adaptation_units(d=256):
  # resnet_v1.bottleneck(..., scope='l1_features') (resnet50_extended_model_hierarchical.py:59-72): an explicit scope REPLACES
  # the default 'bottleneck_v1' of tf.variable_scope(scope, 'bottleneck_v1') - the variables are adaptation_module/<branch>/convK/...
  # (confirmed by running the reference's model() over tests/golden/tf_shim, which records the names it asks for)
  return [UnitSpec(f'adaptation_module/{br}', d, d, d, 1, 1, False) for br, _ in BRANCHES]


PSP_SCOPES = tuple('feature_extractor/pyramid_module/Conv' + ('' if i == 0 else f'_{i}') for i in range(5))
PSP_BINS = (1, 2, 3, 6)


FOV_SCOPE = 'feature_extractor/extension/increase_fov'
# --upsampling_method hybrid: one 3x3 slim.conv2d_transpose (stride 1, SAME, + bias, NO normaliser: the arg scope
# only configures slim.conv2d) per head before the bilinear resize, default scopes inside
# variable_scope('upsampling') (models/resnet50_extended_model_hierarchical.py:84-86,161-179)
UPSAMPLING_SCOPES = tuple('softmax_classifier/upsampling/Conv2d_transpose' + ('' if i == 0 else f'_{i}') for i in range(3))


def conv_specs(head_widths, output_stride=8, d=256, psp=False, fov=None, upsampling='bilinear'):
  """All convolutions in parameter-arena order.  The three adaptation conv1 kernels are adjacent
  so that they form one [3*d, 1, 1, d] filter bank (one GEMM over the shared input).  psp=True adds the
  five 1x1 convolutions of the pyramid module (slim's default scopes Conv .. Conv_4 under
  feature_extractor/pyramid_module, resnet50_extended_model_hierarchical.py:55-57,186-207).  fov=(kernel
  size, rate) adds the dilated `increase_fov` convolution (--fov_expansion_kernel_size / _rate,
  resnet50_extended_feature_extractor.py:44-49)."""
  specs = [ConvSpec(f'{RES}/conv1', 7, 7, 3, 64, 2, 1, True)]
  for u in units(output_stride):
    if u.has_shortcut_conv:
      specs.append(ConvSpec(f'{u.scope}/shortcut', 1, 1, u.depth_in, u.depth, u.stride, 1, False))
    specs.append(ConvSpec(f'{u.scope}/conv1', 1, 1, u.depth_in, u.bottleneck, 1, 1, True))
    specs.append(ConvSpec(f'{u.scope}/conv2', 3, 3, u.bottleneck, u.bottleneck, u.stride, u.rate, True))
    specs.append(ConvSpec(f'{u.scope}/conv3', 1, 1, u.bottleneck, u.depth, 1, 1, False))
  specs.append(ConvSpec('feature_extractor/extension/decrease_fdims', 1, 1, 2048, d, 1, 1, True))
  if fov:
    specs.append(ConvSpec(FOV_SCOPE, fov[0], fov[0], d, d, 1, fov[1], True))
  if psp:
    for sc in PSP_SCOPES[:4]:
      specs.append(ConvSpec(sc, 1, 1, d, d, 1, 1, True))
    specs.append(ConvSpec(PSP_SCOPES[4], 1, 1, 5 * d, d, 1, 1, True))
  au = adaptation_units(d)
  for u in au:
    specs.append(ConvSpec(f'{u.scope}/conv1', 1, 1, d, d, 1, 1, True))
  for u in au:
    specs.append(ConvSpec(f'{u.scope}/conv2', 3, 3, d, d, 1, 1, True))
    specs.append(ConvSpec(f'{u.scope}/conv3', 1, 1, d, d, 1, 1, False))
  for (_, lg), c in zip(BRANCHES, head_widths):
    specs.append(ConvSpec(f'softmax_classifier/{lg}', 1, 1, d, c, 1, 1, False))
  if upsampling == 'hybrid':
    # stored as the equivalent stride-1 correlation kernel (KRSC, rotated 180 degrees); bias in the beta slot
    for sc, c in zip(UPSAMPLING_SCOPES, head_widths):
      specs.append(ConvSpec(sc, 3, 3, c, c, 1, 1, False))
  return specs


def same_pad_before(k, stride, dilation, in_size):
  """Leading zero padding and output size of a slim conv: `conv2d_same` for k > 1 (explicit
  symmetric-first padding when strided), TF 'SAME' otherwise."""
  k_eff = k + (k - 1) * (dilation - 1)
  if stride == 1:
    return (k_eff - 1) // 2, in_size
  # conv2d_same: pad (k_eff-1)//2 before, the rest after, then VALID
  out = (in_size + (k_eff - 1) - k_eff) // stride + 1
  return (k_eff - 1) // 2, out


def conv_flops(specs, H, W, N=1):
  """2*MACs of the forward pass at network input H x W."""
  # spatial size per layer: conv1 -> /2, pool -> /4, block1 last unit -> /8, then constant
  total = 0
  h2, w2 = -(-H // 2), -(-W // 2)
  h4, w4 = -(-h2 // 2), -(-w2 // 2)
  h8, w8 = -(-h4 // 2), -(-w4 // 2)
  for s in specs:
    if s.scope.endswith('resnet_v1_50/conv1'):
      p, q = h2, w2
    elif '/block1/' in s.scope:
      last = '/unit_3/' in s.scope
      if last and (s.scope.endswith('conv2') or s.scope.endswith('conv3')):
        p, q = h8, w8
      else:
        p, q = h4, w4
    else:
      p, q = h8, w8
    total += 2 * N * p * q * s.R * s.S * s.C * s.K
  return total
