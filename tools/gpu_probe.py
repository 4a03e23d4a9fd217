"""Quick device-side timing probe of the hot kernels at BASELINE sizes (not a bench: no JSON contract).
Writes gpurun_out/probe.log.  Each section is independent (try/except) so one failure does not hide the rest."""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from ess_b200 import ops  # noqa: E402
from ess_b200.e2vid import _interleave  # noqa: E402
from helpers import make_e2vid, make_events, make_labels, make_latents, make_semseg  # noqa: E402

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
LOG = open(os.path.join(ROOT, 'gpurun_out', 'probe.log'), 'a')


def log(*a):
    s = ' '.join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + '\n')
    LOG.flush()


def timeit(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def section(name):
    def deco(f):
        log('==', name)
        try:
            f()
        except Exception:
            log('FAILED', name, traceback.format_exc())
        return f
    return deco


B = int(os.environ.get('PROBE_B', '8'))


@section('lstm tc')
def _():
    m = ess_b200.E2VIDRecurrent.__new__(ess_b200.E2VIDRecurrent)
    for (C, H, W) in ((64, 220, 320), (128, 110, 160), (256, 55, 80)):
        w = torch.randn(4 * C, 2 * C, 3, 3, device='cuda') * 0.03
        b = torch.randn(4 * C, device='cuda') * 0.1
        hi, lo, kinp = ops.pack_weight_tc(w, interleave=4)
        e = dict(lstm_tc=dict(hi=hi, lo=lo, k_per_tap=kinp), lstm_b=_interleave(b, 4))
        x = torch.randn(B, H, W, C, device='cuda')
        hp = torch.randn(B, H, W, C, device='cuda')
        cp = torch.randn(B, H, W, C, device='cuda')
        xp = ops.split_bf16(ops.Seg(x), B, H, W)
        hpp = ops.split_bf16(ops.Seg(hp), B, H, W)
        flops = 2.0 * B * H * W * 4 * C * 2 * C * 9
        for passes in (3, 1):
            t = timeit(lambda: ess_b200.E2VIDRecurrent._lstm_tc(m, e, xp, hpp, cp, B, H, W, C, passes))
            log('lstm_tc C=%d %dx%d B=%d passes=%d: %.3f ms  %.1f TFLOP/s (algorithmic)' % (C, H, W, B, passes, t, flops / t / 1e9))
        from ess_b200._lib import EPI_LSTM
        wp = ops.pack_weight(w, interleave=4)
        t = timeit(lambda: ops.conv([ops.Seg(x), ops.Seg(hp)], wp, e['lstm_b'], B, H, W, H, W, 4 * C, ops.taps_conv(3, 1),
                                    epilogue=EPI_LSTM, aux0=cp), warm=1, iters=2)
        log('lstm_fp32 C=%d: %.3f ms  %.1f TFLOP/s' % (C, t, flops / t / 1e9))


@section('e2vid window')
def _():
    T, C, H, W = 2, 5, 440, 640
    data = make_events(B, T, C, H, W).cuda()
    for mode in ('bf16x3', 'bf16', 'fp32'):
        m = make_e2vid(mode=mode).cuda()
        rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
        for wi in (False, True):
            def step():
                rec.update_reconstruction(data[:, :C], with_image=wi)
            t = timeit(step, warm=2, iters=3)
            log('e2vid window mode=%s with_image=%s B=%d: %.2f ms' % (mode, wi, B, t))


@section('semseg fwd+bwd')
def _():
    K, H, W = 11, 440, 640
    dec = make_semseg(K).cuda()
    lat = make_latents(B, H, W, device='cuda')
    labels = make_labels(B, H, W, K).cuda()
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)

    def fwd():
        with torch.no_grad():
            dec(lat)

    def fwdbwd():
        for p in dec.parameters():
            p.grad = None
        loss = crit(dec(lat)[1], labels)
        loss.backward()
    log('semseg fwd B=%d: %.2f ms' % (B, timeit(fwd, 1, 3)))
    log('semseg fwd+loss+bwd B=%d: %.2f ms' % (B, timeit(fwdbwd, 1, 3)))
    log('max mem GB', torch.cuda.max_memory_allocated() / 2 ** 30)
