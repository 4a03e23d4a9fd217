"""Kernel-level trace of ONE bench step with torch.profiler (CUPTI): per-kernel device time, launch counts and the
device idle time between kernels (launch gaps).  Cheaper than the ncu launch list (no serialisation, seconds).
Usage: python tools/step_trace.py [out.json]"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from ess_b200.optim import RAdam  # noqa: E402
from helpers import make_e2vid, make_events, make_labels, make_semseg  # noqa: E402

B, T, C, H, W, K = 8, 20, 5, 440, 640, 11
MODE = os.environ.get('ESS_B200_MODE', 'f16f8')
e2vid = make_e2vid(mode=MODE).cuda()
dec = make_semseg(K).cuda()
crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
rec = ess_b200.ImageReconstructor(e2vid, H, W, C, 'cuda')
opt = RAdam(dec.parameters(), lr=5e-4, betas=(0., 0.999))
data = make_events(B, T, C, H, W).cuda()
labels = make_labels(B, H, W, K).cuda()


def step():
    for p in dec.parameters():
        p.grad = None
    lat = rec.unroll(data, T, C)[2]
    loss = crit(dec({k: v.detach() for k, v in lat.items()})[1], labels)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
tot = collections.OrderedDict()
busy, gap, last_end, first = 0.0, 0.0, None, None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    first = s if first is None else first
    if last_end is not None and s > last_end:
        gap += s - last_end
    last_end = t if last_end is None else max(last_end, t)
    busy += t - s
    name = e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:48]
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += (t - s) / 1e3
span = (last_end - first) / 1e3
print('mode', MODE)
print('one step: %d kernels, span %.2f ms, kernel time %.2f ms, idle gaps %.2f ms' % (len(ev), span, busy / 1e3, gap / 1e3))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-50s %4d %8.3f ms %5.1f%%' % (k, v[0], v[1], 100 * v[1] / span))
print('--- per-launch durations (ms) of the tcgen05 kernels after the encoder unroll, in order:')
seq = [(e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:28], (e.time_range.end - e.time_range.start) / 1e3)
       for e in ev]
last_lstm = max(i for i, (n, _) in enumerate(seq) if n.startswith('conv_tc_kernel<1>'))
print(' '.join('%s:%.3f' % (n.replace('conv_tc_', '').replace('_kernel', ''), d) for n, d in seq[last_lstm + 1:]
               if n.startswith('conv_tc') or n.startswith('wgrad_tc_k')))
win = [(n, d) for n, d in seq[:last_lstm + 1] if n.startswith('conv_tc')]
print('--- last encoder window:', ' '.join('%s:%.3f' % (n.replace('conv_tc_', '').replace('_kernel', ''), d) for n, d in win[-7:]))
if len(sys.argv) > 1:
    json.dump(dict(span_ms=span, kernel_ms=busy / 1e3, idle_ms=gap / 1e3, kernels={k: v for k, v in tot.items()}),
              open(sys.argv[1], 'w'), indent=1)
