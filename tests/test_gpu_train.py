"""Parity of the TRAINING path (forward with batch statistics -> hierarchical masked loss ->
backward through the whole network) with the oracle's autograd on the same
inputs and weights.

Tolerances
  fp32 check mode (direct fp32 convolutions): losses 1e-4 relative; every gradient tensor cosine
      >= 0.9999 and global relative L2 <= 1e-3 (fp32 accumulation-order noise through ~60 batch-norm
      layers whose batch statistics come from only N*h*w = 128 samples per channel)
  bf16 product path (tcgen05 convolutions, bf16 activations): losses 2e-2 relative (north star);
      gradient cosine >= 0.98 per large tensor and >= 0.99 over the whole gradient arena
"""

import pytest
import torch

from oracle import losses as olosses
from oracle import network as onet
from oracle import weak_labels as oweak

pytestmark = pytest.mark.gpu


def _oracle_step(tf_params, dataset, images, labels):
  params = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in tf_params.items()}
  net = onet.Net(params, dataset, training=True)
  pred = net.forward(images)
  losses = olosses.define_losses(pred, labels, dataset)
  losses['total'].backward()
  grads = {k: v.grad for k, v in params.items() if v.requires_grad}
  return losses, grads, net.new_moving, pred


def _labels(dataset, n_strong, n_bbox, n_image, H, W, seed):
  g = torch.Generator().manual_seed(seed)
  ncls = 20 if dataset == 'cityscapes' else 66
  lab = {'prolabels_per_pixel': torch.randint(0, ncls, (n_strong, H // 8, W // 8), generator=g, dtype=torch.int32)
         .repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()}
  if n_bbox:
    boxes = []
    for _ in range(n_bbox):
      k = int(torch.randint(1, 6, (1,), generator=g))
      cids = torch.randint(0, 14, (k,), generator=g).tolist()
      c = torch.rand(k, 4, generator=g)
      boxes.append([(cid, float(min(a, b)), float(max(a, b)), float(min(cc, d)), float(max(cc, d)))
                    for cid, (a, b, cc, d) in zip(cids, c.tolist())])
    lab['prolabels_per_bbox'] = torch.stack([torch.from_numpy(oweak.bbox_labels(b, H, W)) for b in boxes])
  if n_image:
    lab['prolabels_per_image'] = torch.stack([torch.from_numpy(oweak.image_labels([int(torch.randint(0, 14, (1,), generator=g))], H, W))
                                              for _ in range(n_image)])
  return lab


def _compare(net, params, grads, tag):
  """-> (min cosine over tensors with >= 4096 elements, global cosine, global rel-L2)."""
  got_all, ref_all = [], []
  worst = (1.0, None)
  for s in params.specs:
    ref = grads[f'{s.scope}/weights'].permute(3, 0, 1, 2).reshape(-1)  # HWIO -> KRSC
    o = params.w_off[s.scope]
    got = net.ws.grads[o:o + ref.numel()].cpu()
    got_all.append(got)
    ref_all.append(ref)
    c = s_off = params.c_off[s.scope]
    del c
    for name, base in (('gamma', params.n_conv_pad), ('beta', params.n_conv_pad + params.n_chan_pad)):
      r = grads[f'{s.scope}/BatchNorm/{name}']
      gg = net.ws.grads[base + s_off:base + s_off + s.K].cpu()
      got_all.append(gg)
      ref_all.append(r)
    if ref.numel() >= 4096:
      cos = float(torch.nn.functional.cosine_similarity(got, ref, dim=0))
      if cos < worst[0]:
        worst = (cos, s.scope)
  ga, ra = torch.cat(got_all), torch.cat(ref_all)
  gcos = float(torch.nn.functional.cosine_similarity(ga, ra, dim=0))
  rel = float((ga - ra).norm() / ra.norm())
  print(f'{tag}: worst per-tensor cosine {worst[0]:.6f} ({worst[1]}), global cosine {gcos:.6f}, rel-L2 {rel:.3e}')
  return worst[0], gcos, rel


def _run(cuda, dataset, dtype, n_strong, n_bbox, n_image, H, W, seed):
  from wlseg import hierarchy, network, problem_defs
  hier = hierarchy.Hierarchy(dataset, problem_defs.GENERATORS[dataset]()['cids2labels'])
  tf_params = onet.init_params(dataset, seed=seed, randomize_bn=True, tame=True)
  params = network.Params(hier, cuda)
  params.load_tf_dict(tf_params)
  net = network.TrainNetwork(params, dtype=dtype)
  g = torch.Generator().manual_seed(seed + 7)
  N = n_strong + n_bbox + n_image
  images = torch.rand(N, H, W, 3, generator=g) * 2 - 1
  labels = _labels(dataset, n_strong, n_bbox, n_image, H, W, seed + 9)
  dev_labels = {k: v.to(cuda) for k, v in labels.items()}
  logits = net.forward_train(images.to(cuda))
  losses, dlogits = net.loss_and_grad(logits, dev_labels, H, W)
  net.backward(dlogits)
  torch.cuda.synchronize()
  ref_losses, ref_grads, ref_moving, ref_pred = _oracle_step(tf_params, dataset, images, labels)
  return hier, params, net, logits, losses.cpu(), ref_losses, ref_grads, ref_moving, ref_pred


def test_train_step_fp32_check_mode(cuda):
  hier, params, net, logits, losses, rl, rg, rmov, rpred = _run(cuda, 'cityscapes', torch.float32, 2, 0, 0, 64, 64, 3)
  ref_low = torch.cat(rpred['lowres_logits'], -1)
  got_low = logits[..., :hier.total_channels].cpu()
  assert float((got_low - ref_low.detach()).abs().max()) <= 1e-3 * float(ref_low.abs().max())
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  print('losses', losses.tolist(), want.tolist())
  assert torch.allclose(losses, want, rtol=1e-4, atol=1e-5)
  worst, gcos, rel = _compare(net, params, rg, 'fp32')
  assert worst >= 0.9999 and gcos >= 0.9999 and rel <= 1e-2
  # moving statistics after the step (decay 0.9, unbiased variance)
  for scope in ('feature_extractor/base/resnet_v1_50/conv1', 'feature_extractor/extension/decrease_fdims',
                'softmax_classifier/l1_logits'):
    mm = params.moving_mean(scope).cpu()
    mv = params.moving_var(scope).cpu()
    assert torch.allclose(mm, rmov[f'{scope}/BatchNorm/moving_mean'].detach(), rtol=1e-3, atol=1e-4), scope
    assert torch.allclose(mv, rmov[f'{scope}/BatchNorm/moving_variance'].detach(), rtol=1e-3, atol=1e-4), scope


def test_train_step_fp32_weak_labels(cuda):
  """Mixed strong + bbox + image-level batch (BASELINE configs[3] shape class, small)."""
  hier, params, net, logits, losses, rl, rg, _, _ = _run(cuda, 'cityscapes', torch.float32, 1, 1, 1, 48, 64, 5)
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  print('losses', losses.tolist(), want.tolist())
  assert torch.allclose(losses, want, rtol=1e-4, atol=1e-5)
  worst, gcos, rel = _compare(net, params, rg, 'fp32 weak')
  assert worst >= 0.9999 and gcos >= 0.9999 and rel <= 1e-2


@pytest.mark.parametrize('dataset', ['cityscapes', 'vistas'])
def test_train_step_bf16(cuda, dataset):
  hier, params, net, logits, losses, rl, rg, _, rpred = _run(cuda, dataset, torch.bfloat16, 2, 0, 0, 64, 96, 11)
  ref_low = torch.cat(rpred['lowres_logits'], -1).detach()
  got_low = logits[..., :hier.total_channels].cpu()
  el2 = float((got_low - ref_low).norm() / ref_low.norm())
  print(f'bf16 train logits rel-L2 {el2:.3e}')
  assert el2 <= 2e-2
  want = torch.stack([rl['l1_segmentation'], rl['l2_vehicle_segmentation'], rl['l2_human_segmentation'],
                      rl['segmentation']]).detach()
  print('losses', losses.tolist(), want.tolist())
  assert torch.allclose(losses, want, rtol=2e-2, atol=2e-3)
  worst, gcos, rel = _compare(net, params, rg, f'bf16 {dataset}')
  assert worst >= 0.98 and gcos >= 0.99
