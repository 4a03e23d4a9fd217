#!/usr/bin/env python
"""bench.py -- headline benchmark of the ESS hot path on B200 (contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--mode bf16x3|bf16|fp32]

Metric (BASELINE.json): samples/s of one supervised training iteration -- frozen E2VID encoder unrolled
over T event windows (forward) + SemSegE2VID decoder forward + Dice/CE loss + backward (+ RAdam step)
-- on synthetic [B, T*C, H, W] voxel grids.  Workload at every N: BASELINE.json configs[2] "DSEC shape
640x440, 5 bins, 11 classes, batch=8, ess_supervised" per GPU (weak scaling: 8 samples per GPU; for
N > 1 the minibatch is sharded by sample with global-batch semantics, ess_b200/dp.py).

One JSON line on rank 0.  `value` = device-resident inputs; `e2e` = the same step through the public
module API with HOST (pinned) inputs: H2D copy of the events + labels and D2H read of the loss inside
the timed region.  `--impl reference` times the CPU restatement of the reference (oracle/, "port":
the reference is pure Python/PyTorch and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E2VID_CFG = dict(num_bins=5, skip_type='sum', recurrent_block_type='convlstm', num_encoders=3, base_num_channels=32,
                 num_residual_blocks=2, norm='BN', use_upsample_conv=False)
WORK = dict(B=8, T=20, C=5, H=440, W=640, K=11)
METRIC = 'samples/sec fwd+bwd 640x440x5bin voxel grids'


# stdout must carry exactly ONE JSON line (the driver parses it): keep a private handle on the real stdout
# and send everything else any library prints (NCCL banners, torchrun notices, warnings) to stderr.
_REAL_STDOUT = os.fdopen(os.dup(1), 'w')
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + '\n')
    _REAL_STDOUT.flush()


def flops_per_sample(T, C, H, W, K, contract='B'):
    """Algorithmic FLOPs (2*MAC) of one sample, SURVEY.md s8d / BASELINE.md s3."""
    P = H * W
    f_enc = (1600 * C + 519168) * P
    f_img = 150592 * P
    f_seg = (331776 + 64 * K) * P
    if contract == 'A':
        return T * (f_enc + f_img) + 3 * f_seg
    return T * f_enc + f_img + 3 * f_seg


def synth_inputs(B, T, C, H, W, K, seed):
    g = torch.Generator().manual_seed(seed)
    data = torch.randn(B, T * C, H, W, generator=g) * (torch.rand(B, T * C, H, W, generator=g) < 0.2)
    labels = torch.randint(0, K, (B, H, W), generator=g)
    labels[:, :5] = 255
    return data, labels


def randomize_bn_(module, seed=7):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16=p.get('bf16_tflops_sustained', 1401.6), bf16_burst=p.get('bf16_tflops', 1645.8),
                    hbm=p.get('hbm_gbs', 6549.8), source='measured')
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, source='fallback')


def bind_to_gpu_numa_node(index):
    """Restrict this process to the CPUs NVML reports as local to GPU `index` (no-op when NVML or the affinity
    call is unavailable).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return '%d of %d cpus' % (len(cpus), ncpu)
    except Exception as ex:      # diagnostics only; never take the measurement down
        return 'unavailable (%s)' % type(ex).__name__
    return 'unavailable'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        os.remove(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------ CPU reference arm
CPU_SAMPLE_B = 2     # samples of the bounded CPU run (full T = 20 windows each, no extrapolation)
CPU_SAMPLE_NOTE = ('B=%d samples of the workload, all T=20 event windows (image decoder on the last one) + SemSeg decoder '
                   'forward/backward + loss at 440x640 through oracle/ess_oracle.py (the reference\'s own PyTorch CPU ops), '
                   'all host threads' % CPU_SAMPLE_B)


def cpu_reference_sample(threads=None, batch=CPU_SAMPLE_B):
    """Times the oracle (CPU restatement of the reference path: the same PyTorch/MKL-DNN operators the reference
    calls) on a bounded sample of the workload: `batch` samples through one complete supervised iteration -- the
    T-window unroll, image decoder on the last window, decoder forward, Dice+CE, backward.  No extrapolation.
    Returns (samples/s, parts)."""
    from oracle import ess_oracle as O
    w = WORK
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(6)
    import ess_b200
    m = ess_b200.E2VIDRecurrent(dict(E2VID_CFG), mode='fp32')      # parameter container only (CPU tensors)
    randomize_bn_(m)
    dec = ess_b200.SemSegE2VID(256, w['K'], skip_connect=True, skip_type='concat')
    e_sd = {k: v.detach() for k, v in m.state_dict().items()}
    d_sd = {k: v.detach() for k, v in dec.state_dict().items()}
    data, labels = synth_inputs(batch, w['T'], w['C'], w['H'], w['W'], w['K'], 1234)
    t0 = time.perf_counter()
    with torch.no_grad():
        _, _, latent = O.encoder_unroll(e_sd, E2VID_CFG, data, w['T'], w['C'], with_image_last=True)
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    params = {k: v.clone().requires_grad_(True) for k, v in d_sd.items()}
    pred = O.semseg_forward(params, {k: v.detach() for k, v in latent.items()})
    loss = O.task_loss(pred[1], labels, w['K'])
    torch.autograd.grad(loss, list(params.values()))
    t_dec = time.perf_counter() - t0
    return batch / (t_enc + t_dec), dict(batch=batch, t_unroll_s=t_enc, t_decoder_fwd_bwd_s=t_dec)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vals = []
    for i in range(args.warmup + args.steps):
        v, parts = cpu_reference_sample()
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    sample = CPU_SAMPLE_NOTE
    line = dict(impl='reference', metric=METRIC, value=value, unit='samples/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1000.0 / value, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload='DSEC 440x640, C=5 bins, T=20 windows, K=11, ess_supervised (contract B: image on '
                                     'last window)', batch_per_gpu=WORK['B'], parallelism='cpu'),
                cpu_baseline=dict(value=value, unit='samples/s', cores=torch.get_num_threads(), kind='port',
                                  sample=sample, parts=parts),
                e2e=dict(value=value, unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)
    return 0


# ----------------------------------------------------------------------- optional workload: UDA iteration
def run_uda(args):
    """`--workload uda`: BASELINE.json configs[3] -- one `ESSModel.train_step` (training/ess_trainer.py:103-148, DSEC
    branch: image-encoder task step, T-window event unroll, cycle + task-consistency losses, two backward passes
    with the decoder frozen for the first, two RAdam steps) assembled from the drop-in modules, B images + B event
    stacks per step.  Single GPU; NOT the headline metric (bench.py's default workload is the supervised step) --
    unit is (image, event-stack) pairs per second.  Same timing rules: device time over K steps after W warm-ups,
    inputs alternate between two resident batches (> L2); e2e = inputs from pinned host memory + loss read back."""
    import ess_b200
    from ess_b200 import _lib
    from ess_b200.optim import RAdam
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product arm')
    if int(os.environ.get('WORLD_SIZE', '1')) != 1:
        raise RuntimeError('--workload uda is a single-GPU line')
    dev = torch.device('cuda', 0)
    _lib.check(_lib.lib().essb_device_check(), 'device check')
    w = dict(WORK, B=args.batch, T=args.windows)
    B, T, C, H, W, K = w['B'], w['T'], w['C'], w['H'], w['W'], w['K']
    torch.manual_seed(6)
    e2vid = ess_b200.E2VIDRecurrent(dict(E2VID_CFG), mode=args.mode)
    randomize_bn_(e2vid)
    e2vid = e2vid.to(dev).eval()
    for p in e2vid.parameters():
        p.requires_grad = False
    torch.manual_seed(3)
    enc = ess_b200.StyleEncoderE2VID(1, skip_connect=True).to(dev).train()
    dec = ess_b200.SemSegE2VID(256, K, skip_connect=True, skip_type='concat').to(dev)
    rec = ess_b200.ImageReconstructor(e2vid, H, W, C, dev)
    task = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    l1, js = ess_b200.L1Loss(), ess_b200.symJSDivLoss()
    opt_f = RAdam(enc.parameters(), lr=5e-4, betas=(0., 0.999))
    opt_b = RAdam(dec.parameters(), lr=5e-4, betas=(0., 0.999))

    def make(seed):
        g = torch.Generator().manual_seed(seed)
        img = torch.rand(B, 1, H, W, generator=g)
        data, labels = synth_inputs(B, T, C, H, W, K, seed)
        return img.pin_memory(), labels.pin_memory(), data.pin_memory()

    host = [make(1234 + i) for i in range(2)]
    devb = [tuple(t.to(dev) for t in h) for h in host]

    def step(img_a, labels_a, data_b):
        opt_f.zero_grad()
        opt_b.zero_grad()
        lat_fake = enc(img_a)                                                      # ess_trainer.py:150-180
        t_img = task(dec({k: v.detach() for k, v in lat_fake.items()})[1], labels_a)
        t_img.backward()
        img_fake, _, lat_real = rec.unroll(data_b, T, C)                           # :277-280
        lat_real = {k: v.detach() for k, v in lat_real.items()}
        lat_fake = enc(img_fake.detach())                                          # :282
        e_loss = l1(lat_fake[2], lat_real[2]) + l1(lat_fake[4], lat_real[4]) + l1(lat_fake[8], lat_real[8])
        pred_second = dec(lat_fake)                                                # :211-255
        with torch.no_grad():
            pred_first_ng = dec(lat_real)
        e_loss = e_loss + js(pred_second[1], pred_first_ng[1]) + l1(pred_second[2], pred_first_ng[2]) + \
            l1(pred_second[4], pred_first_ng[4])
        pred_first = dec(lat_real)                                                 # :303-330
        with torch.no_grad():
            pred_second_ng = dec({k: v.detach() for k, v in lat_fake.items()})
        t_loss = js(pred_first[1], pred_second_ng[1]) + l1(pred_first[2], pred_second_ng[2]) + \
            l1(pred_first[4], pred_second_ng[4])
        for p in dec.parameters():                                                 # :133-137
            p.requires_grad = False
        e_loss.backward()
        for p in dec.parameters():
            p.requires_grad = True
        t_loss.backward()                                                          # :138
        opt_f.step()
        opt_b.step()
        return t_img.detach() + e_loss.detach() + t_loss.detach()

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for i in range(max(args.warmup, 3)):
        step(*devb[i & 1])
    clocks = ClockSampler(0)
    clocks.start()
    l0 = _lib.launch_count
    ms_dev = timed(lambda i: step(*devb[i & 1]), args.steps)
    launches = _lib.launch_count - l0

    def e2e_step(i):
        h = host[i & 1]
        loss = step(*(t.to(dev, non_blocking=True) for t in h))                    # H2D of the step's inputs
        return float(loss.item())                                                  # D2H of its result

    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    clk = clocks.stop()
    value = args.steps * B / (ms_dev / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    line = dict(metric='(image, event-stack) pairs/sec of one UDA iteration, 640x440x5bin voxel grids', value=value,
                unit='pairs/s', n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_dev / args.steps,
                higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype={'bf16x3': 'bf16x3-split (f32 accumulate, f32 epilogues)', 'bf16': 'bf16 (f32 accumulate)',
                       'fp32': 'f32'}[args.mode],
                data='synthetic',
                config=dict(workload='DSEC 440x640 UDA (ESSModel.train_step, DSEC branch): %d images + %d event stacks, '
                                     'C=5 bins, T=%d windows, K=11' % (B, B, T), batch_per_gpu=B, mode=args.mode,
                            l2_policy='two alternating input batches (larger than the 126 MB L2)'),
                e2e=dict(value=args.steps * B / (ms_e2e / 1e3), unit='pairs/s', ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=h2d, d2h_bytes_per_step=4),
                gpu_launches=launches,
                clocks=dict(sm_mhz=clk['sm_mhz'], sm_max_mhz=clk['sm_max_mhz'], reasons=clk['reasons'], samples=clk['samples']),
                roofline=None, cpu_baseline=None)
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default=os.environ.get('ESS_B200_MODE', 'bf16x3'), choices=['bf16x3', 'bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=WORK['B'], help='samples per GPU')
    ap.add_argument('--windows', type=int, default=WORK['T'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-optimizer', action='store_true')
    ap.add_argument('--profile-all', action='store_true',
                    help='bracket every tcgen05 launch (not only the ConvLSTM cell) with CUDA events inside the timed region')
    ap.add_argument('--workload', default='supervised', choices=['supervised', 'uda'],
                    help="'supervised' (default, the BASELINE.json metric) or 'uda' (configs[3]: one ESSModel.train_step; single GPU)")
    ap.add_argument('--torch-gpu-baseline', action='store_true',
                    help='also time the oracle (the reference op stream as plain PyTorch/cuDNN ops, TF32 default) on this GPU')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload == 'uda':
        return run_uda(args)
    if args.warmup < 3:
        args.warmup = 3

    import ess_b200
    from ess_b200 import _lib, dp
    from ess_b200.optim import RAdam
    # stdout carries exactly ONE JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    rank, world, local = dp.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product arm')
    dev = torch.device('cuda', local)
    _lib.check(_lib.lib().essb_device_check(), 'device check')
    w = dict(WORK, B=args.batch, T=args.windows)
    B, T, C, H, W, K = w['B'], w['T'], w['C'], w['H'], w['W'], w['K']

    torch.manual_seed(6)
    e2vid = ess_b200.E2VIDRecurrent(dict(E2VID_CFG), mode=args.mode)
    randomize_bn_(e2vid)
    e2vid = e2vid.to(dev).eval()
    for p in e2vid.parameters():
        p.requires_grad = False
    dec = ess_b200.SemSegE2VID(256, K, skip_connect=True, skip_type='concat').to(dev)
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], gamma=2.0, num_classes=K, ignore_index=255)
    rec = ess_b200.ImageReconstructor(e2vid, H, W, C, dev)
    bucket = None
    if world > 1:
        dp.attach(rec, crit)
        bucket = dp.GradBucket(dec.parameters())
    opt = None if args.no_optimizer else RAdam([p for p in dec.parameters() if p.requires_grad], lr=5e-4,
                                               weight_decay=0., betas=(0., 0.999))

    # two distinct device-resident batches (901 MB each >> 126 MB L2) alternate between timed steps
    host = [synth_inputs(B, T, C, H, W, K, 1234 + 17 * rank + i) for i in range(2)]
    # Pinned staging memory is allocated from threads bound to the GPU's own NUMA node (what `numactl
    # --cpunodebind` does for a DataLoader process): on a two-socket host a far-node buffer halves the H2D rate,
    # and at 70 ms per step the 919 MB of inputs need > 13 GB/s to stay hidden behind the compute.
    affinity_all = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
    numa = bind_to_gpu_numa_node(local)
    host = [(d.pin_memory(), l.pin_memory()) for d, l in host]
    devb = [(d.to(dev), l.to(dev)) for d, l in host]
    stage_d = torch.empty_like(devb[0][0])
    stage_l = torch.empty_like(devb[0][1])

    def step(data, labels):
        if bucket is not None:
            bucket.zero_()
        else:
            for p in dec.parameters():
                p.grad = None
        _, _, latent = rec.unroll(data, T, C)                     # ess_supervised_trainer.py:126-130
        latent = {k: v.detach() for k, v in latent.items()}       # :145-146
        pred = dec(latent)                                        # :148
        loss = crit(pred[1], labels)                              # :149
        loss.backward()                                           # :103
        if bucket is not None:
            bucket.allreduce_()
        if opt is not None:
            opt.step()                                            # :106
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms)

    for i in range(args.warmup):
        step(*devb[i & 1])
    clocks = ClockSampler(local)
    clocks.start()
    # Inside the timed region only the dominant kernel (fused ConvLSTM cell) is bracketed with CUDA events on its
    # launching stream (60 launches/step); bracketing all ~220 tcgen05 launches costs ~3 % of the step, so the
    # other roles are timed in an extra, untimed pass afterwards (--profile-all puts them back inside).
    _lib.PROFILE = []
    _lib.PROFILE_TAGS = None if args.profile_all else {'lstm_tc'}
    launches0 = _lib.launch_count
    ms_dev = timed(lambda i: step(*devb[i & 1]), args.steps)
    launches = _lib.launch_count - launches0
    prof = _lib.PROFILE
    if not args.profile_all:
        _lib.PROFILE, _lib.PROFILE_TAGS = [], None
        for i in range(2):
            step(*devb[i & 1])
        torch.cuda.synchronize()
        prof_other = [(t, fl, a, b, 2) for (t, fl, a, b) in _lib.PROFILE if t != 'lstm_tc']
    else:
        prof_other = []
    prof = [(t, fl, a, b, args.steps) for (t, fl, a, b) in prof] + prof_other
    _lib.PROFILE, _lib.PROFILE_TAGS = None, None

    # End-to-end: every step's events + labels come from pinned HOST memory and its loss is read back to the
    # host, all inside the timed region.  The H2D copy of step i+1 is issued on a copy stream while step i
    # computes (double-buffered staging), the way a DataLoader(pin_memory) + non_blocking .to() pipeline runs.
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [(stage_d, stage_l), (torch.empty_like(stage_d), torch.empty_like(stage_l))]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        d, l = host[i & 1]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i & 1])          # the compute that last read this buffer is done
            stage[i & 1][0].copy_(d, non_blocking=True)
            stage[i & 1][1].copy_(l, non_blocking=True)
            ready[i & 1].record(copy_stream)

    def e2e_run(steps):
        for ev in freed:
            ev.record()
        prefetch(0)
        losses = []
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i & 1])
            loss = step(*stage[i & 1])
            freed[i & 1].record()
            losses.append(float(loss.item()))               # D2H read of the step's result
        return losses

    # plain H2D rate of the staging copy (diagnostic: explains an e2e value that falls below `value`)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    stage_d.copy_(host[0][0], non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = stage_d.numel() * 4 / (h0.elapsed_time(h1) / 1e3) / 1e9
    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_e2e = float(t)
    clk = clocks.stop()

    samples = args.steps * B * world
    value = samples / (ms_dev / 1e3)
    e2e_value = samples / (ms_e2e / 1e3)
    peaks = load_peaks()

    # roofline of the dominant kernel (the fused tcgen05 ConvLSTM cell): algorithmic FLOPs per launch /
    # mean launch duration from CUDA events recorded on the launching stream inside the timed region
    roof = None
    by_tag = {}
    for tag, fl, a, b, nsteps in prof:
        t = by_tag.setdefault(tag, [0.0, 0.0, 0, nsteps])
        t[0] += fl
        t[1] += a.elapsed_time(b) / 1e3
        t[2] += 1
    if 'lstm_tc' in by_tag:
        fl, sec, n, _ = by_tag['lstm_tc']
        ach = fl / sec / 1e12
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tpath) and (B, T, H, W) == (WORK['B'], WORK['T'], WORK['H'], WORK['W']):
            traffic = json.load(open(tpath)).get('lstm_tc', {}).get('mean_dram_bytes_per_launch')
        roof = dict(kernel='conv_tc_kernel<LSTM> (fused ConvLSTM cell, %d launches)' % n, bound='tensor', achieved=ach,
                    peak=peaks['bf16'], unit='TFLOP/s', frac=ach / peaks['bf16'], traffic=traffic,
                    traffic_note='mean DRAM bytes per launch from the committed ncu --set full capture '
                                 '(profiles/ncu_traffic.json); algorithmic bytes are 865/433/216 MB for the 3 levels',
                    peak_source='%s bf16 dense sustained (MEASURED_PEAKS.json)' % peaks['source'],
                    note='algorithmic FLOPs (2*MAC, no credit for the 3 split passes): in bf16x3 mode the tensor pipe '
                         'executes 3x this, so the mode ceiling is peak/3',
                    mean_launch_ms=sec / n * 1e3, share_of_step=sec * 1e3 / ms_dev,
                    mma_frac_of_peak=ach * (3 if args.mode == 'bf16x3' else 1) / peaks['bf16'])
        # the other tcgen05 launches of the step, by role (algorithmic TFLOP/s, share of the timed region)
        roof['other_tc_kernels'] = {
            tag: dict(achieved=fl2 / sec2 / 1e12, mean_launch_ms=sec2 / n2 * 1e3, launches_per_step=n2 // ns2,
                      share_of_step=(sec2 * 1e3 / ns2) / (ms_dev / args.steps))
            for tag, (fl2, sec2, n2, ns2) in sorted(by_tag.items()) if tag != 'lstm_tc'}
    elif args.mode == 'fp32':
        roof = dict(kernel='conv_fp32_kernel', bound='tensor', achieved=None, peak=peaks['bf16'], unit='TFLOP/s',
                    frac=None, traffic=None)

    line = dict(metric=METRIC, value=value, unit='samples/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype={'bf16x3': 'bf16x3-split (f32 accumulate, f32 epilogues)', 'bf16': 'bf16 (f32 accumulate)',
                       'fp32': 'f32'}[args.mode],
                data='synthetic',
                config=dict(workload='DSEC 440x640, C=5 bins, T=%d windows, K=11, ess_supervised (contract B: E2VID image '
                                     'decoder on the last window only)' % T,
                            batch_per_gpu=B, global_batch=B * world, parallelism='dp%d' % world, mode=args.mode,
                            optimizer='none' if opt is None else 'RAdam (fused kernel)',
                            l2_policy='two alternating 901 MB input batches (inputs larger than the 126 MB L2)',
                            gflop_per_sample=flops_per_sample(T, C, H, W, K) / 1e9),
                tflops=value * flops_per_sample(T, C, H, W, K) / 1e12,
                e2e=dict(value=e2e_value, unit='samples/s', ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=stage_d.numel() * 4 + stage_l.numel() * 8, d2h_bytes_per_step=4,
                         h2d_gb_per_s=h2d_gbs, host_numa_binding=numa),
                gpu_launches=launches, clocks=dict(sm_mhz=clk['sm_mhz'], sm_max_mhz=clk['sm_max_mhz'],
                                                   reasons=clk['reasons'], samples=clk['samples']),
                roofline=roof)
    if affinity_all is not None:
        try:
            os.sched_setaffinity(0, affinity_all)       # the CPU baseline gets every host core back
        except Exception:
            pass
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            v, parts = cpu_reference_sample(threads=cores)
            line['cpu_baseline'] = dict(value=v, unit='samples/s', cores=torch.get_num_threads(), kind='port',
                                        sample=CPU_SAMPLE_NOTE,
                                        parts=parts)
        except Exception as ex:   # the baseline must never take the measurement down
            line['cpu_baseline'] = dict(value=None, error=repr(ex))
    if rank == 0 and world == 1 and args.torch_gpu_baseline:
        # Secondary baseline of SURVEY.md s8d: the reference's op stream through PyTorch/ATen/cuDNN on this B200
        # (cudnn.allow_tf32 = True, torch's default): the oracle's functions run on CUDA tensors.  Bounded
        # sample: B=8, 3 windows (+ image decoder once) + decoder fwd/bwd, extrapolated linearly in T.
        try:
            from oracle import ess_oracle as O
            e_sd = {k: v.detach() for k, v in e2vid.state_dict().items()}
            d_sd = {k: v.detach() for k, v in dec.state_dict().items()}
            d0, l0 = devb[0]

            def enc_windows(nw, with_img):
                st = None
                for i in range(nw):
                    _, st, lat = O.reconstructor_step(e_sd, E2VID_CFG, d0[:, i * C:(i + 1) * C], st,
                                                      with_image=(with_img and i == nw - 1))
                return lat

            def dec_step(lat):
                params = {k: v.clone().requires_grad_(True) for k, v in d_sd.items()}
                pred = O.semseg_forward(params, {k: v.detach() for k, v in lat.items()})
                torch.autograd.grad(O.task_loss(pred[1], l0, K), list(params.values()))

            def tm(fn, n=3):
                fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / n

            with torch.no_grad():
                t3 = tm(lambda: enc_windows(3, False))
                t1 = tm(lambda: enc_windows(1, False))
                t3i = tm(lambda: enc_windows(3, True))
                lat = enc_windows(2, False)
            t_win = (t3 - t1) / 2.0
            t_dec = tm(lambda: dec_step(lat))
            ms = T * t_win + (t3i - t3) + t_dec
            line['torch_gpu_baseline'] = dict(value=B / (ms / 1e3), unit='samples/s', ms_per_step=ms,
                                              kind='oracle ops (PyTorch/ATen/cuDNN, allow_tf32 default) on this GPU',
                                              parts=dict(ms_window=t_win, ms_image_decoder=t3i - t3, ms_decoder_fwd_bwd=t_dec),
                                              sample='B=%d: 3 windows + image decoder once + decoder fwd/bwd, linear in T' % B)
        except Exception as ex:
            line['torch_gpu_baseline'] = dict(value=None, error=repr(ex))
    if rank == 0:
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
