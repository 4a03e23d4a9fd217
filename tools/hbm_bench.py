"""Achieved HBM bandwidth of every memory-bound kernel of the step at the shapes of BASELINE.json configs[2]
(B=8, DSEC 440x640, K=11), timed in place with CUDA events (no profiler):

    python tools/hbm_bench.py [out.md]

Each kernel runs REPS times over ROTATING operand sets whose combined footprint exceeds the 126 MB L2 (so a launch never
finds its inputs in L2 -- colder than in the real step, where a producer's output often still is), bracketed by events on
the launching stream.  bytes = the kernel's ALGORITHMIC traffic (compulsory reads + writes of its tensors); peak =
MEASURED_PEAKS.json hbm_gbs (fallback 6449.1).  `per step` = launches of that shape in one training step.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ess_b200  # noqa: E402,F401
from ess_b200 import ops  # noqa: E402
from ess_b200.ops import Seg  # noqa: E402

dev = torch.device('cuda')
B, T, C, H, W, K = 8, 20, 5, 440, 640, 11
PEAK = 6449.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
REPS = 12
rows = []


def bench(name, per_step, nbytes, make, run, sets=None):
    """make() -> one operand set; run(set) launches the kernel(s)."""
    per_set = max(nbytes, 1)
    n_sets = sets or max(2, min(8, int(300e6 // per_set) + 2))
    S = [make() for _ in range(n_sets)]
    for s in S:
        run(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):
        run(S[i % n_sets])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    gbs = nbytes / (ms / 1e3) / 1e9
    rows.append((name, per_step, nbytes / 1e6, ms * 1e3, gbs, gbs / PEAK))
    print('%-58s x%-3d %8.1f MB %8.1f us %7.0f GB/s  %.2f' % (name, per_step, nbytes / 1e6, ms * 1e3, gbs, gbs / PEAK), flush=True)
    del S
    torch.cuda.empty_cache()


def rnd(*shape):
    return torch.randn(*shape, device=dev)


# ---- event pre-processing
data = rnd(B, T * C, H, W) * (torch.rand(B, T * C, H, W, device=dev) < 0.2)
bench('event_stats (all T windows, one launch)', 1, data.numel() * 4, lambda: data, lambda s: ops.event_stats(s, T, C), sets=2)
stats = ops.event_stats(data, T, C)
cpad = 8


def mk_planes():
    return ops.head_planes_alloc(B, H, W, cpad, dev)


bench('event_prepare_planes (one window -> bf16 hi/lo planes)', T, B * C * H * W * 4 + 2 * B * H * W * cpad * 2, mk_planes,
      lambda s: ops.event_prepare_planes(data[:, :C], stats[0], True, H, W, 0, 0, s))
del data

# ---- InstanceNorm family of the SemSegE2VID decoder (shape, launches per step)
SHAPES = [((55, 80, 256), 11), ((110, 160, 128), 2), ((220, 320, 64), 3), ((440, 640, 32), 1)]
for (h, w, c), cnt in SHAPES:
    n = B * h * w * c
    bench('in_stats [%dx%dx%d]' % (h, w, c), cnt, n * 4, lambda: rnd(B, h, w, c), lambda s: ops.in_stats(s))
for (h, w, c), cnt in SHAPES[:3]:
    n = B * h * w * c

    def mk():
        y = rnd(B, h, w, c)
        m, r = ops.in_stats(y)
        return y, m, r, rnd(B, h, w, c)
    bench('norm_act_add (IN + ReLU + residual) [%dx%dx%d]' % (h, w, c), 5 if c == 256 else 1, n * 12, mk,
          lambda s: ops.norm_act_add(s[0], s[1], s[2], relu=True, res=s[3]))
for (h, w, c), cnt in SHAPES:
    n = B * h * w * c

    def mk():
        y = rnd(B, h, w, c)
        m, r = ops.in_stats(y)
        return y, m, r
    bench('split (IN + ReLU -> operand planes) [%dx%dx%d]' % (h, w, c), cnt, n * 8, mk,
          lambda s: ops.split_bf16(Seg(s[0], mean=s[1], rstd=s[2], relu=True), B, h, w))
for (h, w, c) in [(220, 320, 64)]:
    n = B * h * w * c

    def mk():
        y = rnd(B, h, w, c)
        m, r = ops.in_stats(y)
        return y, m, r
    bench('split (IN + ReLU + nearest x2 -> planes) [%dx%dx%d -> x2]' % (h, w, c), 1, n * 4 + 4 * n * 4, mk,
          lambda s: ops.split_bf16(Seg(s[0], ups=1, mean=s[1], rstd=s[2], relu=True), B, 2 * h, 2 * w))
for (h, w, c), cnt in SHAPES:
    n = B * h * w * c
    ld = (c + 63) // 64 * 64

    def mk():
        y = rnd(B, h, w, c)
        m, r = ops.in_stats(y)
        return y, m, r, rnd(B, h, w, c)
    # pass 1 reads dA + y, writes g; pass 2 reads g + y, writes the gradient as operand planes (4 B / element)
    bench('in_backward -> planes (pass1 + reduce + pass2) [%dx%dx%d]' % (h, w, c), cnt,
          n * 4 * 3 + n * 4 * 2 + B * h * w * ld * 4, mk,
          lambda s: ops.in_backward(s[3], s[0], s[1], s[2], relu=True, planes_ld=ld, want_fp32=False))

# ---- 1x1 classifier (32 -> K) at full resolution and the loss
P = B * H * W


def mk_pw():
    x = rnd(B, H, W, 32)
    m, r = ops.in_stats(x)
    return x, m, r, rnd(K, 32), rnd(K), rnd(B, H, W, K)


bench('pw_conv_fwd (IN + ReLU on load, 32 -> 11)', 1, P * (32 + K) * 4, mk_pw,
      lambda s: ops.pw_conv_fwd(Seg(s[0], mean=s[1], rstd=s[2], relu=True), s[3], s[4], B, H, W, K))
bench('pw_conv_dgrad (11 -> 32)', 1, P * (32 + K) * 4, mk_pw, lambda s: ops.pw_conv_dgrad(s[5], s[3], 32))
bench('pw_conv_wgrad (+ bias gradient)', 1, P * (32 + K) * 4, mk_pw,
      lambda s: ops.pw_conv_wgrad(Seg(s[0], mean=s[1], rstd=s[2], relu=True), s[5]))


def mk_loss():
    lab = torch.randint(0, K, (B, H, W), device=dev)
    lab[:, :5] = 255
    return rnd(B, H, W, K), lab


bench('task_loss_fwd (softmax + CE + Dice partials)', 1, P * (K * 4 + 8), mk_loss,
      lambda s: ops.task_loss_sums(s[0], s[1], K, 255))
sums = ops.task_loss_sums(*mk_loss(), K, 255)
gs = torch.ones(1, device=dev)
bench('task_loss_bwd', 1, P * (2 * K * 4 + 8), mk_loss, lambda s: ops.task_loss_bwd(s[0], s[1], K, 255, sums, True, True, gs))
bench('upsample2_bwd [440x640x64 -> 220x320]', 1, B * 440 * 640 * 64 * 4 * 1.25, lambda: rnd(B, 440, 640, 64),
      lambda s: ops.upsample2_bwd(s, 220, 320, 64))

# ---- optimizer: one multi-tensor launch over the decoder's 6.69 M parameters (28 B / parameter)
from ess_b200.optim import RAdam  # noqa: E402
dec = ess_b200.SemSegE2VID(256, K, skip_connect=True, skip_type='concat').to(dev)
opt = RAdam(dec.parameters(), lr=5e-4, betas=(0., 0.999))
for p in dec.parameters():
    p.grad = torch.randn_like(p)
npar = sum(p.numel() for p in dec.parameters())
bench('radam_multi (34 tensors, one launch)', 1, npar * 28, lambda: None, lambda s: opt.step(), sets=2)

if len(sys.argv) > 1:
    with open(sys.argv[1], 'w') as f:
        f.write('# Achieved HBM bandwidth of the memory-bound kernels, timed in place (tools/hbm_bench.py, CUDA events, rotating\n'
                '# operand sets larger than L2; B=8, DSEC 440x640, K=11).  Peak = %.1f GB/s (MEASURED_PEAKS.json hbm_gbs).\n\n' % PEAK)
        f.write('| kernel [shape] | launches / step | algorithmic MB | us / launch | GB/s | frac of peak | ms / step |\n|---|---|---|---|---|---|---|\n')
        tot = 0.0
        for name, per, mb, us, gbs, fr in rows:
            f.write('| %s | %d | %.1f | %.1f | %.0f | %.2f | %.3f |\n' % (name, per, mb, us, gbs, fr, per * us / 1e3))
            tot += per * us / 1e3
        f.write('\nSum over the listed launches: %.2f ms per step.\n' % tot)
