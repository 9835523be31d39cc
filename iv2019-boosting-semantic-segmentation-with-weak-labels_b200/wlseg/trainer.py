"""One data-parallel training step: the host-side mirror of the TRAIN branch of `define_estimator`
(code/estimator/define_estimator_hierarchical.py:77-159), `define_optimizer`
(code/estimator/define_optimizer.py:3-26) and the MirroredStrategy gradient exchange the reference
gets from `RunConfig(train_distribute=...)` (code/system_factory.py:279-295).

  forward (batch statistics) -> fused hierarchical loss fwd+bwd -> backward through the network
  -> gradient all-reduce (mean over replicas of each replica's own normalised-loss gradient, the
     [TF-1.12] MirroredStrategy semantics) overlapped with backward
  -> fused SGD-momentum + L2 + bf16 operand refresh (one launch over the parameter arena)
  -> optional EMA shadows (forced off under --distribute, system_factory.py:236-238)

Launch sequencing and buffer ownership only; the arithmetic is in libwlseg.
"""

import os

import torch

from wlseg import network, ops


class GradientBuckets:
  """Reverse-order bucketing of a flat gradient arena for overlap with backward.

  The arena is laid out in FORWARD layer order, backward finishes its tail first.  `ready(lo)` tells
  the bucketer that every element at offset >= lo is final; whole buckets that became final are
  all-reduced asynchronously on `comm_stream` (NCCL) or inline (gloo / CPU tests); `finish()`
  flushes the rest and averages.  Device agnostic so that the logic is testable with gloo."""

  def __init__(self, flat, bucket_elems, world_size, process_group=None, comm_stream=None, extra=(), tail_elems=0):
    self.flat = flat
    self.extra = list(extra)  # tensors that only become final at the very end (BN gamma / beta gradients)
    self.also_wait = []       # further streams whose work a bucket depends on (the wgrad side stream)
    self.n = flat.numel()
    self.world = world_size
    self.group = process_group
    self.stream = comm_stream
    # bucket boundaries from the tail: [n - k*b, n - (k-1)*b).  The LOWEST addresses (root convolution, block1)
    # become final last, and whatever bucket holds them is exposed in front of the optimizer: it is cut down to
    # `tail_elems` (measured at 2 GPUs: a 19 MB last bucket = 75 us exposed, profiles/r2_timeline_train_n2.txt)
    self.bounds = []
    hi = self.n
    tail = min(tail_elems, self.n) if tail_elems else 0
    while hi > tail:
      lo = max(tail, hi - bucket_elems)
      self.bounds.append((lo, hi))
      hi = lo
    if tail:
      self.bounds.append((0, tail))
    self.next = 0
    self.handles = []
    self.launched_bytes = 0

  def reset(self):
    self.next = 0
    self.handles = []

  def _launch(self, lo, hi):
    self._launch_view(self.flat[lo:hi])

  def _launch_view(self, view):
    import torch.distributed as dist
    self.launched_bytes += view.numel() * view.element_size()
    if self.world == 1:
      return
    if self.stream is not None:
      self.stream.wait_stream(torch.cuda.current_stream())
      for st in self.also_wait:
        self.stream.wait_stream(st)
      with torch.cuda.stream(self.stream):
        dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
    else:
      self.handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

  def ready(self, lo):
    while self.next < len(self.bounds) and self.bounds[self.next][0] >= lo:
      self._launch(*self.bounds[self.next])
      self.next += 1

  def finish(self):
    """All remaining buckets; returns the factor the caller must scale the summed gradient by
    (1/world: mean over replicas) - folded into the optimizer kernel's grad_scale."""
    if self.world > 1 and self.stream is not None and self.next < len(self.bounds) and self.extra:
      # the last bucket and the BN parameter gradients leave as ONE grouped collective (one NCCL kernel)
      import torch.distributed as dist
      views = [self.flat[lo:hi] for lo, hi in self.bounds[self.next:]] + list(self.extra)
      self.next = len(self.bounds)
      self.stream.wait_stream(torch.cuda.current_stream())
      for st in self.also_wait:
        self.stream.wait_stream(st)
      with torch.cuda.stream(self.stream):
        with dist._coalescing_manager(group=self.group, device=self.flat.device):
          for v in views:
            self.launched_bytes += v.numel() * v.element_size()
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
    else:
      self.ready(0)
      for t in self.extra:
        self._launch_view(t)
    for h in self.handles:
      h.wait()
    self.handles = []
    if self.stream is not None and self.world > 1:
      torch.cuda.current_stream().wait_stream(self.stream)
    return 1.0 / self.world


class Trainer:
  def __init__(self, params, settings, dtype=torch.bfloat16, rank=0, world_size=1, bucket_mb=25, use_graph=True):
    """use_graph: after two eager steps (lazy initialisation, kernel attributes, allocator warm-up)
    the whole step - ~520 kernel launches - is captured into ONE CUDA graph per input signature and
    replayed; the learning rate is read from device memory so the graph stays valid across the
    schedule.  Inputs are copied into static buffers before each replay."""
    self.use_graph = bool(use_graph) and params.device.type == 'cuda'
    self._graphs = {}
    self._eager_steps = 0
    self.p = params
    self.s = settings
    self.rank, self.world = rank, world_size
    # --cross_replica_norm (models/resnet50_extended_model_hierarchical.py:257,327-328): BN moments over all replicas
    xr = (world_size, None) if (getattr(settings, 'cross_replica_norm', False) and world_size > 1) else None
    self.net = network.TrainNetwork(params, dtype=dtype, bn_decay=getattr(settings, 'batch_norm_decay', 0.9),
                                    cross_replica=xr)
    self.ws = self.net.ws
    self.momentum = float(getattr(settings, 'momentum', 0.9))
    self.nesterov = bool(getattr(settings, 'use_nesterov', False))
    self.optimizer = getattr(settings, 'optimizer', 'SGDM')
    if self.optimizer not in ('SGD', 'SGDM'):
      raise ValueError('Optimizer is not supported.')  # define_optimizer.py:24
    self.wd = float(getattr(settings, 'regularization_weight', 0.00017))
    self.ema_decay = 0.0 if getattr(settings, 'distribute', False) else float(getattr(settings, 'ema_decay', 0.0))
    if self.ema_decay > 0:
      # [TF-1.12] ExponentialMovingAverage.apply on tf.Variables: the shadow slot starts at the
      # variable's initial value and zero-debiasing is NOT applied (it only applies to plain tensors)
      self.ws.ema_shadow = params.master.clone()
    self.comm_stream = torch.cuda.Stream(device=params.device) if (world_size > 1 and params.device.type == 'cuda') else None
    # (WLSEG_CONV_SMS=144 + NCCL_MAX_CTAS=4 would give the collectives SMs of their own next to the persistent
    # convolution kernels - csrc/common.cuh conv_sms; measured at 8 GPUs it does NOT pay: 11.77 ms/step with all 148 SMs
    # and NCCL's defaults, 11.86 with 144 + 4, 12.54 with 140 + 8, profiles/r2_n8_sm_partition.txt)
    # last (exposed) bucket: everything below block2 - conv1 + block1, ~0.9 MB - so that the tail of the exchange is
    # one small collective; it travels with the BN gamma / beta gradients, which also become final at the very end
    spec_off = [params.w_off[s.scope] for s in params.specs if '/block2/' in s.scope]
    tail = min(spec_off) if spec_off else 0
    self.buckets = GradientBuckets(self.ws.grads[:params.n_conv_pad], bucket_mb * (1 << 20) // 4, world_size,
                                   comm_stream=self.comm_stream, extra=[self.ws.grads[params.n_conv_pad:]],
                                   tail_elems=tail)
    self.net.grad_ready = self.buckets.ready if world_size > 1 else None
    if os.environ.get('WLSEG_WGRAD_STREAM', '0') == '1' and params.device.type == 'cuda':
      self.net.wgrad_stream = torch.cuda.Stream(device=params.device)
      self.buckets.also_wait.append(self.net.wgrad_stream)   # a bucket is final only when its wgrads have run
    self.grad_payload = 'fp32'
    self.global_step = 0
    self._lr_host = None

  def set_lr(self, lr):
    if lr != self._lr_host:
      self.ws.lr.fill_(float(lr))
      self._lr_host = lr

  def _core(self, images, labels):
    """forward + loss + backward + gradient exchange + optimizer; returns the device loss vector."""
    net, ws, p = self.net, self.ws, self.p
    H, W = images.shape[1], images.shape[2]
    self.buckets.reset()
    logits = net.forward_train(images)
    losses, dlogits = net.loss_and_grad(logits, labels, H, W)
    net.backward(dlogits)
    gscale = self.buckets.finish()
    ws.reg_loss.zero_()
    mom = self.momentum if self.optimizer == 'SGDM' else 0.0
    ops.sgdm_step(p.master, ws.grads, ws.momentum, p.operand, p.n_conv_pad, ws.lr, mom, self.nesterov, self.wd, gscale,
                  ws.reg_loss)
    p.touch()
    out = torch.empty(6, dtype=torch.float32, device=p.device)
    reg = ws.reg_loss.to(torch.float32)
    out[0] = losses[3] + reg[0]
    out[1] = losses[3]
    out[2:5] = losses[0:3]
    out[5] = reg[0]
    return out

  def step(self, features, labels, lr):
    """-> device tensor float[6]: total, segmentation, l1, l2_vehicle, l2_human, regularization."""
    images = features['proimages']
    labels = {k: v for k, v in labels.items() if v is not None}
    self.set_lr(lr)
    if self.ema_decay > 0:
      # ema.apply sits in UPDATE_OPS, which create_train_op runs BEFORE the gradient step
      # (define_estimator_hierarchical.py:96-111,120-129): the shadows average the variables as they are before this
      # step's update, with num_updates = the pre-increment global_step
      t = self.global_step
      d = min(self.ema_decay, (1.0 + t) / (10.0 + t))
      ops.ema_update(self.ws.ema_shadow, self.ws.ema_shadow, self.p.master, d, 1.0)
    if not self.use_graph or self.net.profile is not None or self.net.keep:
      out = self._core(images, labels)
    else:
      key = (tuple(images.shape), images.dtype) + tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in labels.items()))
      entry = self._graphs.get(key)
      if entry is None and self._eager_steps < 2:
        self._eager_steps += 1
        out = self._core(images, labels)
      else:
        if entry is None:
          static_img = images.clone()
          static_lab = {k: v.clone() for k, v in labels.items()}
          torch.cuda.synchronize()
          graph = torch.cuda.CUDAGraph()
          with torch.cuda.graph(graph):
            static_out = self._core(static_img, static_lab)
          entry = (graph, static_img, static_lab, static_out)
          self._graphs[key] = entry
          # the capture itself did not execute anything: fall through to the first replay
        graph, static_img, static_lab, static_out = entry
        static_img.copy_(images, non_blocking=True)
        for k, v in labels.items():
          static_lab[k].copy_(v, non_blocking=True)
        graph.replay()
        self.p.touch()
        out = static_out.clone()   # the graph overwrites static_out on the next replay
    self.global_step += 1
    return out
