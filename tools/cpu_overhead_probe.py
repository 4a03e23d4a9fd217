"""How far ahead of the GPU does the host run?  Host-side (no sync) vs device time of the two halves of a step."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from helpers import make_e2vid, make_events, make_labels, make_semseg  # noqa: E402

B, T, C, H, W, K = 8, 20, 5, 440, 640, 11
e2vid = make_e2vid(mode='bf16x3').cuda()
dec = make_semseg(K).cuda()
crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
rec = ess_b200.ImageReconstructor(e2vid, H, W, C, 'cuda')
data = make_events(B, T, C, H, W).cuda()
labels = make_labels(B, H, W, K).cuda()


def enc():
    return rec.unroll(data, T, C)[2]


def decstep(lat):
    for p in dec.parameters():
        p.grad = None
    loss = crit(dec({k: v.detach() for k, v in lat.items()})[1], labels)
    loss.backward()


lat = enc()
decstep(lat)
torch.cuda.synchronize()
for name, fn in (('encoder unroll', lambda: enc()), ('decoder fwd+loss+bwd', lambda: decstep(lat))):
    host, dev = [], []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        host.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        dev.append(e0.elapsed_time(e1))
    print('%-24s host-side issue time %.1f ms   device time %.1f ms' % (name, min(host), min(dev)))
