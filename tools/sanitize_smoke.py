"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
One supervised iteration (fused unroll with flip + hot pixels, decoder fwd/bwd, loss, RAdam) in the f16f8 (default), bf16x3
and fp32 modes with the split-K tail switched on -- B = 2 at 48 x 80 makes the 1/8 level row-stacked (tall view, masked rows)
and runs the CTA-pair, merged-phase and halo kernels -- then the same unroll replayed as a CUDA graph, the N = 128 CTA-pair
variant, a ConvGRU encoder window, the UDA image encoder fwd/bwd."""
import os
import sys
import types

import torch

os.environ.setdefault('ESS_B200_PRETRAINED', '0')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from ess_b200 import ops  # noqa: E402
from ess_b200.optim import RAdam  # noqa: E402
from helpers import E2VID_CFG, make_e2vid, make_events, make_labels, make_semseg  # noqa: E402

ops.SPLITK = True
B, T, C, H, W, K = 2, 2, 5, 48, 80, 6
for mode in ('f16f8', 'bf16x3', 'fp32'):
    e2vid = make_e2vid(mode=mode).cuda()
    dec = make_semseg(K).cuda()
    dec.mode = mode
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    opts = types.SimpleNamespace(flip=True, hot_pixels_file=None, no_normalize=False, no_recurrent=False, color=False)
    rec = ess_b200.ImageReconstructor(e2vid, H, W, C, 'cuda', opts)
    rec.set_hot_pixels([(3, 4), (70, 40)])
    opt = RAdam(dec.parameters(), lr=5e-4, betas=(0., 0.999))
    data, labels = make_events(B, T, C, H, W).cuda(), make_labels(B, H, W, K).cuda()
    for _ in range(2):
        for p in dec.parameters():
            p.grad = None
        img, _, lat = rec.unroll(data, T, C)
        loss = crit(dec({k: v.detach() for k, v in lat.items()})[1], labels)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    print(mode, 'supervised step ok, loss %.5f' % float(loss))
    if mode == 'f16f8':
        for _ in range(2):
            rec.unroll(data, T, C, graph=True)          # capture + replay, then replay only
        torch.cuda.synchronize()
        os.environ['ESSB_TC_PAIR128'] = '2'             # opt-in CTA-pair kernel for the N = 128 tiles
        rec.unroll(data, T, C)
        os.environ['ESSB_TC_PAIR128'] = '0'
        torch.cuda.synchronize()
        print(mode, 'graphed unroll + N=128 pair kernel ok')
# 168 tiles on 148 SMs: exercises the split-K tail of the scheduler
m = ess_b200.E2VIDRecurrent(dict(E2VID_CFG), mode='bf16x3').cuda().eval()
st = None
ev = make_events(1, 2, 5, 224, 384).cuda()
for i in range(2):
    with torch.no_grad():
        _, st, lat = m(ev[:, i * 5:(i + 1) * 5], st, with_image=False)
torch.cuda.synchronize()
print('large-tile LSTM windows ok')
g = ess_b200.E2VIDRecurrent(dict(E2VID_CFG, recurrent_block_type='convgru'), mode='bf16x3').cuda().eval()
st = None
for i in range(2):
    with torch.no_grad():
        _, st, _ = g(ev[:, i * 5:(i + 1) * 5, :48, :80].contiguous(), st)
torch.cuda.synchronize()
print('ConvGRU windows ok')
enc = ess_b200.StyleEncoderE2VID(1, skip_connect=True).cuda().train()
x = torch.rand(2, 1, 48, 80, device='cuda')
out = enc(x)
(out[8].sum() + out[4].sum() + out[2].sum()).backward()
torch.cuda.synchronize()
print('StyleEncoderE2VID fwd/bwd ok')
