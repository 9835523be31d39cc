"""N > 1 host logic on CPU: world_size-2 gloo.

  * GradientBuckets: tail-first bucketed all-reduce of a flat gradient arena == mean over ranks of
    the per-rank gradients (the MirroredStrategy semantics the reference trains with,
    code/system_factory.py:279-283), for ready-notifications in any order
  * evaluation sharded by image: per-rank integer confusion matrices summed == the single-rank
    confusion matrix of all images, bit-exact (code/estimator/define_estimator_hierarchical.py:185-194), through the
    oracle's histogram and through the product's `SemanticSegmentation._reduce_across_ranks`; rank 0's precondition
    verdict reaches every rank (`_any_rank`)
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200')


def _worker(rank, world, port, q):
  for p in (ROOT, PKG):
    if p not in sys.path:
      sys.path.insert(0, p)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from oracle import metrics as ometrics
    from wlseg.trainer import GradientBuckets
    # ---- gradient buckets
    n = 10007
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=g)
    extra = torch.randn(33, generator=g)
    mine = flat.clone()
    b = GradientBuckets(flat, 1024, world, extra=[extra])
    b.reset()
    for lo in (9000, 9500, 4000, 4001, 100):  # non-monotone notifications must be harmless
      b.ready(lo)
    assert b.next == sum(1 for (l, h) in b.bounds if l >= 100)
    scale = b.finish()
    assert b.next == len(b.bounds)
    got = flat * scale
    all_mine = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(all_mine, mine)
    want = torch.stack(all_mine).mean(0)
    ok_grad = bool(torch.allclose(got, want, rtol=1e-6, atol=1e-7))
    # ---- sharded evaluation
    rng = np.random.RandomState(7)
    labels = rng.randint(0, 20, size=(8, 16, 16)).astype(np.int32)
    decs = rng.randint(0, 20, size=(8, 16, 16)).astype(np.int32)
    full = ometrics.confusion_matrix(labels, decs, 20)
    part = ometrics.confusion_matrix(labels[rank::world], decs[rank::world], 20)
    t = torch.from_numpy(part.astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ok_cm = bool(np.array_equal(t.numpy(), full))
    # ---- the product's own N > 1 host logic: SemanticSegmentation._reduce_across_ranks and the collective verdict on
    # rank 0's log-directory precondition (wlseg/system_factory.py)
    import types
    from wlseg import system_factory as sf
    st = types.SimpleNamespace(world_size=world, device='cpu', rank=rank)
    holder = types.SimpleNamespace(_settings=st, _estimator=types.SimpleNamespace(device='cpu'))
    m = sf.SemanticSegmentation._reduce_across_ranks(
        holder, {'confusion_matrix_int64': part.astype(np.int64).copy(), 'confusion_matrix': part.astype(np.int32), 'steps': 4})
    ok_cm = ok_cm and bool(np.array_equal(m['confusion_matrix_int64'], full)) and m['confusion_matrix'].dtype == np.int32 \
        and bool(np.array_equal(m['confusion_matrix'], full))
    ok_cm = ok_cm and sf._any_rank(rank == 0, st) is True and sf._any_rank(False, st) is False
    q.put((rank, ok_grad, ok_cm))
  finally:
    dist.destroy_process_group()


def test_world2_gloo_gradient_buckets_and_sharded_eval():
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=240) for _ in range(world)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, ok_grad, ok_cm in res:
    assert ok_grad, f'rank {rank}: bucketed all-reduce != mean of per-rank gradients'
    assert ok_cm, f'rank {rank}: sharded confusion matrix != single-rank confusion matrix'


def test_get_temp_Nb_splits_the_global_batch():
  """code/input_pipelines/utils.py:118-124."""
  for p in (ROOT, PKG):
    if p not in sys.path:
      sys.path.insert(0, p)
  import types
  import pytest
  from wlseg.estimator import get_temp_Nb
  s = types.SimpleNamespace(distribute=True, world_size=4)
  assert get_temp_Nb(s, 8) == 2
  with pytest.raises(AssertionError):
    get_temp_Nb(s, 6)
  assert get_temp_Nb(types.SimpleNamespace(distribute=False), 6) == 6
