"""In-situ kernel timeline of the replayed training / evaluation step (CUPTI activity records through
torch.profiler: kernels run back to back inside the CUDA graph exactly as in the bench, unlike ncu's serialised
cold-cache replays).  Prints per-kernel totals per step, the busy time and the idle gaps between kernels."""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from wlseg import hierarchy, network, ops, problem_defs, synthetic, trainer as wtrainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'train'
# under torchrun (WORLD_SIZE > 1): data-parallel training step with the bucketed NCCL gradient all-reduce; rank 0 prints
world, rank, local_rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
if world > 1:
  import torch.distributed as dist
  dist.init_process_group('nccl', device_id=dev)
hier = hierarchy.Hierarchy('cityscapes', problem_defs.cityscapes()['cids2labels'])
params = network.Params(hier, dev)
params.init_random(0)
src = synthetic.SyntheticInputs(hier.num_classes, dev, rank=rank)
STEPS = 3
if mode == 'train':
  class S:
    momentum, use_nesterov, optimizer, regularization_weight = 0.9, False, 'SGDM', 0.00017
    batch_norm_decay, distribute, ema_decay = 0.9, world > 1, 0.0
  tr = wtrainer.Trainer(params, S, dtype=torch.bfloat16, rank=rank, world_size=world)
  batches = [src.train_batch(4, 0, 0, 768, 768) for _ in range(2)]

  def step(i):
    f, l = batches[i % 2]
    return tr.step(f, {k: v for k, v in l.items() if v is not None}, 0.01)
else:
  net = network.Network(params, dtype=torch.bfloat16)
  batches = [src.eval_batch(4, 1024, 2048) for _ in range(2)]
  evstep = network.EvalStep(net, 20)   # the product path: one CUDA graph per resident batch

  def step(i):
    f, l = batches[i % 2]
    evstep(f['proimages'], l['prolabels'])
for i in range(5):
  step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  for i in range(STEPS):
    step(i)
  torch.cuda.synchronize()
if world > 1:
  dist.barrier()
if rank != 0:
  if mode == 'train':
    tr._graphs.clear()
  torch.cuda.synchronize()
  dist.destroy_process_group()
  sys.exit(0)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy, span0, span1 = 0.0, ev[0].time_range.start, max(e.time_range.end for e in ev)
agg = collections.OrderedDict()
gaps, last_end = [], None
for e in ev:
  d = e.time_range.end - e.time_range.start
  name = re.sub(r'\(.*', '', e.name)[:70]
  a = agg.setdefault(name, [0, 0.0])
  a[0] += 1
  a[1] += d
  busy += d
  if last_end is not None:
    gaps.append(max(0.0, e.time_range.start - last_end))
  last_end = max(last_end or 0, e.time_range.end)
span = span1 - span0
print(f'{mode}: {STEPS} steps, span {span / STEPS / 1e3:.3f} ms/step, kernel busy {busy / STEPS / 1e3:.3f} ms/step, '
      f'{len(ev) // STEPS} kernels/step, idle between kernels {sum(gaps) / STEPS / 1e3:.3f} ms/step '
      f'(median gap {sorted(gaps)[len(gaps) // 2]:.2f} us)')
print(f'{"us/step":>10} {"n/step":>7} {"share":>6}  kernel')
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
  print(f'{t / STEPS:10.1f} {n / STEPS:7.1f} {100 * t / busy:5.1f}%  {name}')

# ---- collectives: where each NCCL kernel sits in the step and what ran beside it
nccl = [e for e in ev if 'nccl' in e.name.lower()]
if nccl:
  comp = [e for e in ev if 'nccl' not in e.name.lower()]
  print(f'\nNCCL kernels of the LAST profiled step ({len(nccl) // STEPS} per step): start offset in the step, duration, '
        f'compute kernels running while it ran, compute-idle time inside it')
  per = len(nccl) // STEPS
  step_start = None
  last = nccl[-per:]
  first_comp_of_step = min((e.time_range.start for e in comp if e.time_range.start >= last[0].time_range.start - 12e3), default=last[0].time_range.start)
  for e in last:
    s0, s1 = e.time_range.start, e.time_range.end
    over = [c for c in comp if c.time_range.end > s0 and c.time_range.start < s1]
    covered = sum(min(c.time_range.end, s1) - max(c.time_range.start, s0) for c in over)
    names = collections.Counter(re.sub(r'void wlseg::|<.*', '', c.name)[:24] for c in over)
    print(f'  {e.name.split('(')[0][:60]:60s} +{(s0 - first_comp_of_step) / 1e3:7.3f} ms  {(s1 - s0):8.1f} us  '
          f'compute busy {covered:8.1f} us  {dict(names.most_common(3))}')
  # exposed communication: from the end of the last backward kernel to the start of the optimizer
  sg = [c for c in comp if 'sgdm' in c.name]
  if sg:
    opt = sg[-1]
    before = [c for c in comp if c.time_range.end <= opt.time_range.start]
    last_bwd = max(c.time_range.end for c in before)
    print(f'  optimizer starts {(opt.time_range.start - last_bwd):.1f} us after the last compute kernel before it '
          f'(= exposed tail of the gradient exchange + launch gap)')
  if mode == 'train' and world > 1:
    tr._graphs.clear()
    torch.cuda.synchronize()
    dist.destroy_process_group()

# WLSEG_TIMELINE_LIST=<regex>: per-launch durations (mean over the profiled steps) of the matching kernels in step order,
# with the kernel that ran before each one - to match bandwidth passes to layers
pat = os.environ.get('WLSEG_TIMELINE_LIST')
if pat:
  per = len(ev) // STEPS
  rows = collections.OrderedDict()
  for i, e in enumerate(ev):
    if re.search(pat, e.name):
      k = i % per
      r = rows.setdefault(k, [re.sub(r'\(.*', '', e.name)[:44], re.sub(r'\(.*', '', ev[i - 1].name)[:60], 0.0, 0])
      r[2] += e.time_range.end - e.time_range.start
      r[3] += 1
  print(f'---- per-launch list of /{pat}/ ({len(rows)} per step)')
  for k, (n, prev, t, c) in rows.items():
    print(f'{k:5d} {t / c:8.1f} us  {n:44s} after {prev}')
