"""GPU + CPU-oracle probe (not a test): how closely does the bf16 PRODUCT training step follow the oracle end to
end, as a function of the input size?  Prints, per size, the low-res logits error, the loss errors and the gradient
cosines against (a) the fp32 oracle and (b) the oracle with the product's storage roundings made explicit, plus a
3-step Trainer-vs-oracle trajectory.  The numbers it prints set the tolerances of tests/test_gpu_train.py and
tests/test_gpu_baseline_shapes.py.

  python tools/train_parity_probe.py 2x64x96 2x128x128 2x256x256 4x768x768
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'iv2019-boosting-semantic-segmentation-with-weak-labels_b200'))
import torch  # noqa: E402

from oracle import losses as olosses  # noqa: E402
from oracle import network as onet  # noqa: E402
from oracle import optimizer as oopt  # noqa: E402
from wlseg import hierarchy, network, problem_defs, trainer as wtrainer  # noqa: E402

dev = torch.device('cuda:0')
dataset = os.environ.get('DATASET', 'cityscapes')
hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
ncls = hier.num_classes
torch.set_num_threads(os.cpu_count() or 1)


def cos(a, b):
  a, b = a.double().reshape(-1), b.double().reshape(-1)
  return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))


def oracle_step(tf_params, images, labels, storage):
  params = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in tf_params.items()}
  net = onet.Net(params, dataset, training=True, storage=storage)
  pred = net.forward(images)
  losses = olosses.define_losses(pred, labels, dataset)
  losses['total'].backward()
  return losses, {k: v.grad for k, v in params.items() if v.requires_grad}, torch.cat(pred['lowres_logits'], -1).detach()


def grads_vs(params, net, ref):
  got_all, ref_all, worst = [], [], (1.0, None)
  for s in params.specs:
    r = ref[f'{s.scope}/weights'].permute(3, 0, 1, 2).reshape(-1)
    o = params.w_off[s.scope]
    g = net.ws.grads[o:o + r.numel()].cpu()
    got_all.append(g)
    ref_all.append(r)
    if r.numel() >= 4096:
      c = cos(g, r)
      if c < worst[0]:
        worst = (c, s.scope[-40:])
  ga, ra = torch.cat(got_all), torch.cat(ref_all)
  return worst, cos(ga, ra), float((ga.double() - ra.double()).norm() / ra.double().norm())


for spec in sys.argv[1:] or ['2x64x96', '2x128x128']:
  N, H, W = (int(x) for x in spec.split('x'))
  seed = 11
  tf_params = onet.init_params(dataset, seed=seed, randomize_bn=True, tame=True)
  # RES_GAMMA < 1 scales the gamma of every residual-branch output: a random-init train-mode BN network amplifies
  # perturbations ~1.08x per layer (x300 end to end, oracle bf16-storage vs oracle fp32 = 0.8 rel-L2 on the logits);
  # small residual gammas give the conditioning of a trained network
  fac = float(os.environ.get('RES_GAMMA', '1.0'))
  for k in tf_params:
    if k.endswith('conv3/BatchNorm/gamma') and 'bottleneck' in k:
      tf_params[k] = tf_params[k] * fac
  g = torch.Generator().manual_seed(seed + 7)
  images = torch.rand(N, H, W, 3, generator=g) * 2 - 1
  labels = {'prolabels_per_pixel': torch.randint(0, ncls, (N, H // 8, W // 8), generator=g, dtype=torch.int32)
            .repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()}
  params = network.Params(hier, dev)
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=torch.bfloat16)
  logits = net.forward_train(images.to(dev))
  losses, dlogits = net.loss_and_grad(logits, {k: v.to(dev) for k, v in labels.items()}, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  got_low = logits[..., :hier.total_channels].cpu()
  got_losses = losses.cpu()
  for storage in ('fp32', 'bf16'):
    t0 = time.perf_counter()
    rl, rg, rlow = oracle_step(tf_params, images, labels, storage)
    dt = time.perf_counter() - t0
    want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                        rl['segmentation']]).detach()
    lerr = ((got_losses - want).abs() / want.abs().clamp_min(1e-12)).tolist()
    worst, gcos, rel = grads_vs(params, net, rg)
    print(f'[res_gamma {fac}] {spec} vs oracle[{storage}] ({dt:.1f}s CPU): logits max-rel {float((got_low - rlow).abs().max() / rlow.abs().max()):.3e} '
          f'rel-L2 {float((got_low - rlow).norm() / rlow.norm()):.3e} | loss rel err {["%.2e" % e for e in lerr]} | '
          f'grad worst cos {worst[0]:.4f} ({worst[1]}) global cos {gcos:.4f} rel-L2 {rel:.3e}', flush=True)
  del net

  # ---- 3 optimizer steps through the product Trainer vs the oracle (fp32 and bf16 storage) ----
  if N * H * W <= 2 * 256 * 256:
    class S:
      momentum, use_nesterov, optimizer, regularization_weight = 0.9, False, 'SGDM', 0.00017
      batch_norm_decay, distribute, ema_decay = 0.9, False, 0.0
    params = network.Params(hier, dev)
    params.load_tf_dict(tf_params)
    w0 = params.master.clone()
    tr = wtrainer.Trainer(params, S, use_graph=False)
    traj = [tr.step({'proimages': images.to(dev)}, {k: v.to(dev) for k, v in labels.items()}, 0.01).cpu().clone() for _ in range(3)]
    torch.cuda.synchronize()
    for storage in ('fp32', 'bf16'):
      p = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in tf_params.items()}
      acc = {k: torch.zeros_like(v) for k, v in p.items() if v.requires_grad}
      otraj = []
      for _ in range(3):
        for v in p.values():
          v.grad = None
        onet_ = onet.Net(p, dataset, training=True, storage=storage)
        l = olosses.define_losses(onet_.forward(images), labels, dataset)
        l['total'].backward()
        otraj.append(float(l['segmentation']))
        with torch.no_grad():
          for k, v in p.items():
            if v.requires_grad:
              wn, an = oopt.momentum_step(v, v.grad + (0.00017 * v if k.endswith('weights') else 0.0), acc[k], 0.01)
              v.copy_(wn)
              acc[k] = an
          for k, v in onet_.new_moving.items():
            p[k].copy_(v.detach())
      # parameter delta after 3 steps, conv kernels only
      d_got, d_ref = [], []
      for s in params.specs:
        o = params.w_off[s.scope]
        n = s.K * s.R * s.S * s.C
        d_got.append((params.master[o:o + n] - w0[o:o + n]).cpu())
        d_ref.append((p[f'{s.scope}/weights'].detach() - tf_params[f'{s.scope}/weights']).permute(3, 0, 1, 2).reshape(-1))
      dg, dr = torch.cat(d_got), torch.cat(d_ref)
      print(f'[res_gamma {fac}] {spec} 3-step trajectory vs oracle[{storage}]: seg loss got {[round(float(t[1]), 5) for t in traj]} ref '
            f'{[round(x, 5) for x in otraj]} | weight-delta cosine {cos(dg, dr):.4f} rel-L2 '
            f'{float((dg.double() - dr.double()).norm() / dr.double().norm()):.3e}', flush=True)
