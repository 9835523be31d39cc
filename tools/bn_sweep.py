"""A/B timing of the batch-norm kernels over the layer shapes of one 4 x 768 x 768 training step
(tools/gpu_exp.sh runs it on the B200).  Flat-index kernels (WLSEG_BN_FLAT=1) against the row-mapped ones for
a few (rows in flight, CTAs per SM) settings; CUDA events, rotating buffers larger than L2, algorithmic GB/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402
from wlseg import ops  # noqa: E402

SHAPES = [(589824, 64, 1), (147456, 64, 4), (147456, 256, 4), (36864, 128, 8), (36864, 256, 14), (36864, 512, 10),
          (36864, 1024, 7), (36864, 2048, 4)]   # (rows, channels, ~launches of that shape per step)
dev = torch.device('cuda:0')


def bufs(rows, C, n):
  k = max(2, int(300e6 // (rows * C * 2 * n)) + 1)   # rotate through > 2 x L2 worth of tensors
  return [[torch.randn(rows, C, device=dev).to(torch.bfloat16) for _ in range(n)] for _ in range(k)]


def timed(fn, sets, reps=24):
  """GPU time per launch: the launches are captured into one CUDA graph (Python + ctypes launch overhead is
  ~10 us per call and would hide every kernel shorter than that) and the graph is replayed."""
  for i in range(3):
    fn(sets[i % len(sets)])
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    for i in range(reps):
      fn(sets[i % len(sets)])
  g.replay()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(3):
    g.replay()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1e3 / (3 * reps)


def run(kind, rows, C):
  f32 = lambda: torch.rand(C, device=dev) + 0.5
  scale, shift, mean, invstd, gamma = f32(), f32(), f32(), f32(), f32()
  dg, db = torch.rand(C, device=dev, dtype=torch.float64), torch.rand(C, device=dev, dtype=torch.float64)
  if kind == 'apply':
    sets = bufs(rows, C, 2)
    return timed(lambda b: ops.bn_apply(b[0], scale, shift, None, b[1], rows, C, True), sets), 4
  if kind == 'apply_res':
    sets = bufs(rows, C, 3)
    return timed(lambda b: ops.bn_apply(b[0], scale, shift, b[2], b[1], rows, C, True), sets), 6
  if kind == 'apply_res_mask':
    sets = bufs(rows, C, 3)
    mask = torch.empty(rows, C // 8, dtype=torch.uint8, device=dev)
    return timed(lambda b: ops.bn_apply(b[0], scale, shift, b[2], b[1], rows, C, True, mask=mask), sets), 6
  if kind == 'bwd':
    sets = bufs(rows, C, 3)
    return timed(lambda b: ops.bn_bwd_apply(b[0], None, b[1], mean, invstd, gamma, dg, db, rows, C, True, b[2], None,
                                            scale=scale, shift=shift, pitch=C), sets), 6
  if kind == 'reduce':
    sets = bufs(rows, C, 2)
    return timed(lambda b: ops.bn_bwd_reduce(b[0], None, b[1], mean, invstd, rows, C, True, dg, db, scale=scale,
                                             shift=shift, pitch=C), sets), 4
  raise ValueError(kind)


def sweep(kind, configs):
  print(f'== {kind}: us per launch (GB/s algorithmic); last column = weighted us per step')
  print('config'.ljust(28) + ''.join(f'{r}x{c}'.rjust(18) for r, c, _ in SHAPES) + '   step_us')
  for name, env in configs:
    for k in KNOBS:
      os.environ.pop(k, None)
    os.environ.update(env)
    line, tot = name.ljust(28), 0.0
    for rows, C, w in SHAPES:
      us, bpe = run(kind, rows, C)
      tot += us * w
      line += f'{us:8.1f} ({rows * C * bpe / us / 1e3:6.0f})'.rjust(18)
    print(line + f'{tot:10.0f}', flush=True)


KNOBS = ('WLSEG_BN_FLAT', 'WLSEG_BN_APPLY_U', 'WLSEG_BN_APPLY_CTAS', 'WLSEG_BN_BWD_U', 'WLSEG_BN_BWD_CTAS',
         'WLSEG_BN_RED_CTAS', 'WLSEG_BN_RED_SPAN')

if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'mask':
  SHAPES = [(147456, 256, 2), (36864, 256, 4), (36864, 512, 4), (36864, 1024, 6), (36864, 2048, 3)]
  sweep('apply_res', [('rows U=2 ctas=4', {})])
  sweep('apply_res_mask', [('rows U=2 ctas=4', {})])
  sys.exit(0)

if __name__ == '__main__':
  sweep('reduce', [(f'span={sp} ctas={c}', {'WLSEG_BN_RED_SPAN': str(sp), 'WLSEG_BN_RED_CTAS': str(c)})
                   for sp, c in ((0, 2), (0, 1), (0, 4), (256, 2), (256, 4), (512, 2), (128, 2))])
  sweep('apply', [('flat', {'WLSEG_BN_FLAT': '1'})] + [(f'rows U={u} ctas={c}', {'WLSEG_BN_APPLY_U': str(u), 'WLSEG_BN_APPLY_CTAS': str(c)})
                                                     for u, c in ((1, 8), (2, 4), (2, 8), (4, 2), (4, 4))])
  sweep('apply_res', [('flat', {'WLSEG_BN_FLAT': '1'})] + [(f'rows U={u} ctas={c}', {'WLSEG_BN_APPLY_U': str(u), 'WLSEG_BN_APPLY_CTAS': str(c)})
                                                         for u, c in ((1, 8), (2, 4), (4, 2))])
  sweep('bwd', [('flat', {'WLSEG_BN_FLAT': '1'})] + [(f'rows U={u} ctas={c}', {'WLSEG_BN_BWD_U': str(u), 'WLSEG_BN_BWD_CTAS': str(c)})
                                                   for u, c in ((1, 3), (1, 6), (2, 2), (2, 3), (2, 4), (4, 1), (4, 2))])
