"""Per-parameter gradient error table of the SemSeg decoder vs the fp64 oracle (debug aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from helpers import O, make_labels, make_latents, make_semseg  # noqa: E402


def oracle(dec, lat, labels, K, dtype):
    params = {k: v.detach().cpu().to(dtype).clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    lat_c = {k: v.detach().cpu().to(dtype) for k, v in lat.items()}
    pred = O.semseg_forward(params, lat_c)
    loss = O.task_loss(pred[1], labels.cpu(), K)
    return pred, dict(zip(params.keys(), torch.autograd.grad(loss, list(params.values()))))


for (K, H, W, B) in ((6, 40, 56, 2), (6, 64, 96, 2), (6, 40, 56, 1)):
    for mode in ('fp32',):
        dec = make_semseg(K).cuda()
        dec.mode = mode
        lat = make_latents(B, H, W, device='cuda')
        labels = make_labels(B, H, W, K).cuda()
        p64, g64 = oracle(dec, lat, labels, K, torch.float64)
        p32, g32 = oracle(dec, lat, labels, K, torch.float32)
        pred = dec(lat)
        crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
        crit(pred[1], labels).backward()
        print('== K=%d %dx%d B=%d mode=%s' % (K, H, W, B, mode))
        for k in (4, 2, 1):
            e = float((pred[k].detach().cpu().double() - p64[k]).abs().max() / p64[k].abs().max())
            e32 = float((p32[k].double() - p64[k]).abs().max() / p64[k].abs().max())
            print('  out[%d] err ours %.2e ref32 %.2e' % (k, e, e32))
        for n, p in dec.named_parameters():
            if n.endswith('bias'):
                continue
            r = g64[n]
            d = p.grad.cpu().double() - r
            d32 = g32[n].double() - r
            print('  %-34s max ours %.2e ref32 %.2e | L2 ours %.2e ref32 %.2e' % (
                n, float(d.abs().max() / r.abs().max()), float(d32.abs().max() / r.abs().max()),
                float(d.norm() / r.norm()), float(d32.norm() / r.norm())))
