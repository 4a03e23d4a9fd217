"""Times one UDA iteration (ESSModel.train_step shape, training/ess_trainer.py:103-148, DSEC branch) built from the
drop-in modules at BASELINE config 4 size (B=8 images + B=8 event stacks, 440x640, T=20, C=5, K=11).
Not the headline bench (bench.py measures the supervised metric); writes gpurun_out/uda_probe.json."""
import json
import os
import sys

import torch

os.environ.setdefault('ESS_B200_PRETRAINED', '0')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from ess_b200.optim import RAdam  # noqa: E402
from helpers import make_e2vid, make_events, make_labels, make_semseg  # noqa: E402

B = int(os.environ.get('PROBE_B', '8'))
T, C, H, W, K = int(os.environ.get('PROBE_T', '20')), 5, 440, 640, 11
mode = os.environ.get('ESS_B200_MODE', 'f16f8')
dev = 'cuda'
e2vid = make_e2vid(mode=mode).to(dev)
torch.manual_seed(3)
enc = ess_b200.StyleEncoderE2VID(1, skip_connect=True).to(dev).train()
dec = make_semseg(K).to(dev)
rec = ess_b200.ImageReconstructor(e2vid, H, W, C, dev)
task = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
l1, js = ess_b200.L1Loss(), ess_b200.symJSDivLoss()
opt_f = RAdam(enc.parameters(), lr=5e-4, betas=(0., 0.999))
opt_b = RAdam(dec.parameters(), lr=5e-4, betas=(0., 0.999))
g = torch.Generator().manual_seed(0)
img_a = torch.rand(B, 1, H, W, generator=g).to(dev)
labels_a = make_labels(B, H, W, K).to(dev)
data_b = make_events(B, T, C, H, W).to(dev)


def step():
    opt_f.zero_grad()
    opt_b.zero_grad()
    lat_fake = enc(img_a)
    t_img = task(dec({k: v.detach() for k, v in lat_fake.items()})[1], labels_a)
    t_img.backward()
    img_fake, _, lat_real = rec.unroll(data_b, T, C)
    lat_real = {k: v.detach() for k, v in lat_real.items()}
    lat_fake = enc(img_fake.detach())
    e_loss = l1(lat_fake[2], lat_real[2]) + l1(lat_fake[4], lat_real[4]) + l1(lat_fake[8], lat_real[8])
    pred_second = dec(lat_fake)
    with torch.no_grad():
        pred_first_ng = dec(lat_real)
    e_loss = e_loss + js(pred_second[1], pred_first_ng[1]) + l1(pred_second[2], pred_first_ng[2]) + \
        l1(pred_second[4], pred_first_ng[4])
    pred_first = dec(lat_real)
    with torch.no_grad():
        pred_second_ng = dec({k: v.detach() for k, v in lat_fake.items()})
    t_loss = js(pred_first[1], pred_second_ng[1]) + l1(pred_first[2], pred_second_ng[2]) + \
        l1(pred_first[4], pred_second_ng[4])
    for p in dec.parameters():
        p.requires_grad = False
    e_loss.backward()
    for p in dec.parameters():
        p.requires_grad = True
    t_loss.backward()
    opt_f.step()
    opt_b.step()
    return t_img + e_loss.detach() + t_loss.detach()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
out = dict(config='DSEC UDA: B=%d images + B=%d event stacks, 440x640, T=%d, C=5, K=11' % (B, B, T), mode=mode,
           ms_per_step=ms, pairs_per_s=B / (ms / 1e3), loss=float(loss), max_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'uda_probe.json'), 'w'))

if os.environ.get('PROBE_TRACE'):
    # kernel-level trace of one UDA iteration (CUPTI): where the ~80 ms outside the encoder unroll go
    import collections
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = collections.OrderedDict()
    for e in ev:
        name = e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:52]
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += (e.time_range.end - e.time_range.start) / 1e3
    total = sum(v[1] for v in tot.values())
    print('UDA iteration: %d kernels, kernel time %.2f ms' % (len(ev), total))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:32]:
        print('%-54s %4d %8.3f ms %5.1f%%' % (k, v[0], v[1], 100 * v[1] / total))
    for pat in ('wgrad_fp32_kernel', 'conv_fp32_kernel', 'wgrad_tc_kernel'):
        print(pat, ' '.join('%.3f' % ((e.time_range.end - e.time_range.start) / 1e3) for e in ev if pat in e.name))
