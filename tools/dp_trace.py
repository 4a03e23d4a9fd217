"""Where does the data-parallel step time go?  Run under torchrun with N ranks (one per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_trace.py [out.json]

Every rank runs the bench step (BASELINE.json configs[2] per GPU, sharded global batch, overlapped per-stage gradient
exchange) and reports, WITHOUT any barrier inside the measurement:
  * its own step time over K steps (CUDA events around each step) -- the ranks synchronise only through the three
    all-reduces of the step, so the slowest rank's pace is visible as waiting time in the others' NCCL kernels;
  * from a CUPTI trace of one step: total kernel time, time inside ncclDevKernel* (transfer + waiting for the peer),
    idle time between kernels;
  * the same step with the collectives detached (each rank alone, no exchange): the rank's own compute pace.
Rank 0 gathers and prints one table + the straggler share: (max over ranks of the solo step) / (mean solo step) - 1.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from ess_b200 import dp  # noqa: E402
from ess_b200.optim import RAdam  # noqa: E402
from helpers import make_e2vid, make_events, make_labels, make_semseg  # noqa: E402

B, T, C, H, W, K = 8, 20, 5, 440, 640, 11
STEPS = int(os.environ.get('DP_TRACE_STEPS', '8'))
rank, world, local = dp.init_from_env()
dev = torch.device('cuda', local)
e2vid = make_e2vid(mode=ess_b200.e2vid.default_mode()).to(dev)      # helpers default to the fp32 kernels
dec = make_semseg(K).to(dev)
crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
rec = ess_b200.ImageReconstructor(e2vid, H, W, C, dev)
opt = RAdam(dec.parameters(), lr=5e-4, betas=(0., 0.999))
data = make_events(B, T, C, H, W, seed=1234 + rank).to(dev)
labels = make_labels(B, H, W, K, seed=99 + rank).to(dev)
bucket = dp.GradBucket(dec.parameters(), module=dec)


def attach(on):
    if on and world > 1:
        dp.attach(rec, crit)
    else:
        rec.stats_reduce_fn = None
        crit.reduce_fn = None


def step(exchange=True):
    bucket.zero_()
    lat = rec.unroll(data, T, C, graph=True)[2]
    loss = crit(dec({k: v.detach() for k, v in lat.items()})[1], labels)
    loss.backward()
    if exchange:
        bucket.allreduce_()
    opt.step()


def timed_steps(exchange):
    for _ in range(3):
        step(exchange)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    for a, b in ev:
        a.record()
        step(exchange)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return dict(mean=sum(ms) / len(ms), median=ms[len(ms) // 2], min=ms[0], max=ms[-1])


# ---- (1) each rank alone: collectives detached (world-size-1 semantics per rank; no waiting anywhere)
attach(False)
real_world = world
dp_world = dp.world_size
dp.world_size = lambda: 1              # GradBucket._launch / allreduce_sum_ become no-ops
solo = timed_steps(False)
dp.world_size = dp_world
# ---- (2) the real data-parallel step
attach(True)
if world > 1:
    dist.barrier()
together = timed_steps(True)
# ---- (3) CUPTI trace of one data-parallel step
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step(True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
kernel_ms = sum(e.time_range.end - e.time_range.start for e in evs) / 1e3
nccl = [e for e in evs if 'nccl' in e.name.lower()]
nccl_ms = sum(e.time_range.end - e.time_range.start for e in nccl) / 1e3
span = (max(e.time_range.end for e in evs) - evs[0].time_range.start) / 1e3
row = dict(rank=rank, solo_ms=solo, dp_ms=together, trace=dict(span_ms=span, kernel_ms=kernel_ms, nccl_ms=nccl_ms,
                                                                  nccl_kernels=len(nccl),
                                                                  nccl_each_ms=[round((e.time_range.end - e.time_range.start) / 1e3, 3)
                                                                                for e in nccl]))
rows = [None] * world
if world > 1:
    dist.all_gather_object(rows, row)
else:
    rows = [row]
if rank == 0:
    print('world %d, B=%d per GPU, DSEC 440x640, T=%d, default mode, graphed unroll, overlapped gradient exchange; %d steps'
          % (world, B, T, STEPS))
    print('%4s %12s %12s %12s %10s %10s %8s' % ('rank', 'solo ms', 'dp ms', 'dp - solo', 'nccl ms', 'kernels ms', '#nccl'))
    for r in rows:
        print('%4d %12.2f %12.2f %12.2f %10.2f %10.2f %8d' % (r['rank'], r['solo_ms']['median'], r['dp_ms']['median'],
                                                            r['dp_ms']['median'] - r['solo_ms']['median'],
                                                            r['trace']['nccl_ms'], r['trace']['kernel_ms'],
                                                            r['trace']['nccl_kernels']))
    solo_med = [r['solo_ms']['median'] for r in rows]
    dp_med = [r['dp_ms']['median'] for r in rows]
    mean_solo, max_solo = sum(solo_med) / world, max(solo_med)
    print('solo step: mean %.2f ms, slowest rank %.2f ms (straggler share %.1f %%); data-parallel step: %.2f ms'
          % (mean_solo, max_solo, 100 * (max_solo / mean_solo - 1), max(dp_med)))
    print('=> of the %.2f ms above the mean solo step, %.2f ms is the slowest GPU\'s own pace (independent power capping) and '
          '%.2f ms is exchange + synchronisation' % (max(dp_med) - mean_solo, max_solo - mean_solo, max(dp_med) - max_solo))
    print('NCCL kernels of rank 0 in the traced step (ms, in launch order; each includes waiting for the slowest peer):',
          rows[0]['trace']['nccl_each_ms'])
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], 'w'), indent=1)
if world > 1:
    dist.destroy_process_group()
