"""A handful of representative convolution launches for `ncu --set full` (one launch each, after a
warm-up launch): eval shapes 4x128x256 (1024x2048 input) and the training wgrad at 4x96x96.
usage: prof_conv.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402
from wlseg import ops  # noqa: E402

dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bf = torch.bfloat16


def fprop(N, H, W, C, K, R, dil, res=False, stats=False, relu=True):
  x = torch.randn(N, H, W, C, device=dev).to(bf)
  w = (torch.randn(K, R, R, C, device=dev) * (2.0 / (R * R * C)) ** 0.5).to(bf)
  y = torch.empty(N, H, W, K, dtype=bf, device=dev)
  r = torch.randn(N, H, W, K, device=dev).to(bf) if res else None
  pad = dil * (R - 1) // 2
  p = ops.conv_params((N, H, W, C), (K, R, R, C), dilation=dil, pad=(pad, pad), out_hw=(H, W), relu=relu, res=r)
  if stats:
    s1 = torch.zeros(K, dtype=torch.float64, device=dev)
    s2 = torch.zeros(K, dtype=torch.float64, device=dev)
    sc = sh = None
  else:
    s1 = s2 = None
    sc = torch.ones(K, device=dev)
    sh = torch.zeros(K, device=dev)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  for i in range(reps):
    if i == reps - 1:
      e0.record()
    ops.conv2d_fprop(p, x, w, y, sc, sh, r, s1, s2)
  e1.record()
  torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 1e3
  fl = 2.0 * N * H * W * R * R * C * K
  print(f'fprop N{N} {H}x{W} C{C} K{K} R{R} d{dil} res{int(res)} stats{int(stats)}: {us:8.1f} us {fl / us / 1e6:7.1f} TF/s')


def wgrad(N, H, W, C, K, R, dil):
  x = torch.randn(N, H, W, C, device=dev).to(bf)
  dy = torch.randn(N, H, W, K, device=dev).to(bf)
  dw = torch.empty(K, R, R, C, dtype=torch.float32, device=dev)
  pad = dil * (R - 1) // 2
  p = ops.conv_params((N, H, W, C), (K, R, R, C), dilation=dil, pad=(pad, pad), out_hw=(H, W))
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  for i in range(reps):
    if i == reps - 1:
      e0.record()
    ops.conv2d_wgrad(p, x, dy, dw)
  e1.record()
  torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 1e3
  fl = 2.0 * N * H * W * R * R * C * K
  print(f'wgrad N{N} {H}x{W} C{C} K{K} R{R} d{dil}: {us:8.1f} us {fl / us / 1e6:7.1f} TF/s')


MEMBOUND = [
    dict(N=4, H=128, W=256, C=512, K=2048, R=1, dil=1, res=True),    # block4 conv3
    dict(N=4, H=128, W=256, C=256, K=1024, R=1, dil=1, res=True),    # block3 conv3
    dict(N=4, H=128, W=256, C=128, K=512, R=1, dil=1, res=True),     # block2 conv3
    dict(N=4, H=256, W=512, C=64, K=256, R=1, dil=1, res=True),      # block1 conv3
    dict(N=4, H=256, W=512, C=64, K=256, R=1, dil=1),                # block1 shortcut
    dict(N=4, H=128, W=256, C=1024, K=256, R=1, dil=1),              # block3 conv1
    dict(N=4, H=128, W=256, C=256, K=768, R=1, dil=1),               # adaptation conv1 x3
    dict(N=4, H=96, W=96, C=256, K=1024, R=1, dil=1, stats=True, relu=False),   # training block3 conv3
    dict(N=4, H=96, W=96, C=1024, K=256, R=1, dil=1, stats=True, relu=False),   # training block3 conv1
    dict(N=4, H=192, W=192, C=64, K=256, R=1, dil=1, stats=True, relu=False),   # training block1 conv3
]

if len(sys.argv) > 2 and sys.argv[2] == 'sweep':
  for ew in (8,):
    os.environ['WLSEG_EW'] = str(ew)
    print(f'--- WLSEG_EW={ew}')
    for kw in MEMBOUND:
      fprop(**kw)
    fprop(4, 128, 256, 512, 512, 3, 4)
    fprop(4, 128, 256, 2048, 512, 1, 1)
    fprop(4, 96, 96, 512, 512, 3, 4, stats=True, relu=False)
  sys.exit(0)

if len(sys.argv) > 2 and sys.argv[2] == 'stats':
  for na in ('0', '1'):
    os.environ['WLSEG_NO_STAT_ATOMICS'] = na
    print('--- WLSEG_NO_STAT_ATOMICS=' + na)
    for st in (False, True):
      fprop(4, 96, 96, 256, 256, 3, 2, stats=st, relu=False)
      fprop(4, 96, 96, 128, 128, 3, 1, stats=st, relu=False)
      fprop(4, 96, 96, 1024, 256, 1, 1, stats=st, relu=False)
      fprop(4, 96, 96, 256, 1024, 1, 1, stats=st, relu=False)
      fprop(4, 96, 96, 512, 2048, 1, 1, stats=st, relu=False)
      fprop(4, 192, 192, 64, 256, 1, 1, stats=st, relu=False)
  sys.exit(0)

if len(sys.argv) > 2 and sys.argv[2] == 'statone':
  fprop(4, 96, 96, 256, 1024, 1, 1, stats=True, relu=False)
  fprop(4, 96, 96, 256, 1024, 1, 1, stats=False, relu=False)
  sys.exit(0)

if len(sys.argv) > 2 and sys.argv[2] == 'pairprobe':
  fprop(4, 128, 256, 256, 1024, 1, 1, res=True)
  fprop(4, 128, 256, 512, 2048, 1, 1, res=True)
  fprop(4, 128, 256, 256, 768, 1, 1)
  sys.exit(0)

if len(sys.argv) > 2 and sys.argv[2] == 'membound':
  for kw in MEMBOUND:
    fprop(**kw)
  sys.exit(0)

fprop(4, 128, 256, 512, 512, 3, 4)             # block4 conv2 (dilated 3x3): the biggest FLOP share
fprop(4, 128, 256, 256, 256, 3, 2)             # block3 conv2
fprop(4, 128, 256, 2048, 512, 1, 1)            # block4 conv1
fprop(4, 128, 256, 512, 2048, 1, 1, res=True)  # block4 conv3 + shortcut add
fprop(4, 128, 256, 1024, 256, 1, 1)            # block3 conv1
fprop(4, 96, 96, 512, 512, 3, 4, stats=True, relu=False)   # training forward (raw z + BN statistics)
wgrad(4, 96, 96, 512, 512, 3, 4)
wgrad(4, 96, 96, 2048, 512, 1, 1)
