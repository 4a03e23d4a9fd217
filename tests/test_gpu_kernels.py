"""Kernel-level parity tests (`-m gpu`): every C-ABI kernel vs the CPU oracle / plain torch fp32 ops on
identical seeded inputs.  Tolerance: 1e-3 max-norm relative (BASELINE.json north_star) unless noted;
the exact-fp32 kernels are expected (and asserted) to be far tighter."""
import pytest
import torch
import torch.nn.functional as F

from helpers import O, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3        # the contract
TIGHT = 2e-5      # what exact-fp32 kernels actually achieve (accumulation-order noise)


def _ops():
    from ess_b200 import ops
    return ops


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(x):
    return x.permute(0, 3, 1, 2).cpu()


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,stride', [
    (2, 16, 24, 16, 32, 3, 1),
    (1, 55, 80, 64, 128, 3, 1),       # odd DSEC 1/8 extent, Cout > 64 (TN=8 path)
    (2, 20, 12, 8, 64, 5, 2),
    (1, 16, 16, 5, 32, 5, 1),         # C=5 head: scalar-load path
    (2, 9, 7, 32, 11, 1, 1),          # 1x1 classifier, Cout not a multiple of 4
    (1, 24, 40, 20, 72, 3, 1),        # channel tail in the K-step (20 = 16 + 4) and Cout tail
])
def test_conv_fp32_linear(N, H, W, Cin, Cout, k, stride):
    ops = _ops()
    from ess_b200._lib import ACT_RELU
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * 0.1
    b = torch.randn(Cout, generator=g)
    pad = k // 2
    ref = torch.relu(F.conv2d(x, w, b, stride=stride, padding=pad))
    OH, OW = ref.shape[2:]
    wp = ops.pack_weight(w.cuda())
    out, _, _, _ = ops.conv([ops.Seg(nhwc(x))], wp, b.cuda(), N, H, W, OH, OW, Cout, ops.taps_conv(k, pad),
                            stride=stride, act=ACT_RELU)
    assert rel_err(nchw(out), ref) < TIGHT


def test_conv_fp32_fused_loader_and_stats():
    """two segments: [upsample2(relu(IN(y))), skip] -> conv3x3, with IN statistics from the epilogue."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    N, H, W, C0, C1, Cout = 2, 12, 20, 32, 16, 48
    y = torch.randn(N, C0, H // 2, W // 2, generator=g) * 2 + 0.5
    skip = torch.randn(N, C1, H, W, generator=g)
    w = torch.randn(Cout, C0 + C1, 3, 3, generator=g) * 0.05
    b = torch.randn(Cout, generator=g)
    a = torch.relu(F.instance_norm(y, eps=1e-5))
    a = a.repeat_interleave(2, 2).repeat_interleave(2, 3)
    ref = F.conv2d(torch.cat([a, skip], 1), w, b, padding=1)
    mean = y.mean((2, 3)).cuda()
    rstd = (1.0 / torch.sqrt(y.var((2, 3), unbiased=False) + 1e-5)).cuda()
    segs = [ops.Seg(nhwc(y), ups=1, mean=mean.contiguous(), rstd=rstd.contiguous(), relu=True), ops.Seg(nhwc(skip))]
    out, _, st, tiles = ops.conv(segs, ops.pack_weight(w.cuda()), b.cuda(), N, H, W, H, W, Cout, ops.taps_conv(3, 1),
                                 want_stats=True)
    assert rel_err(nchw(out), ref) < TIGHT
    m2, r2 = ops.in_finalize(st, H * W)
    assert rel_err(m2.cpu(), ref.mean((2, 3))) < TIGHT
    assert rel_err(r2.cpu(), 1.0 / torch.sqrt(ref.var((2, 3), unbiased=False) + 1e-5)) < 1e-4


def test_conv_transposed_phases():
    ops = _ops()
    from ess_b200._lib import ACT_RELU
    g = torch.Generator().manual_seed(2)
    N, H, W, Cin, Cout = 2, 7, 10, 32, 16
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cin, Cout, 5, 5, generator=g) * 0.1
    skip = torch.randn(N, Cout, 2 * H, 2 * W, generator=g)
    ref = torch.relu(F.conv_transpose2d(x, w, None, stride=2, padding=2, output_padding=1)) + skip
    wp = ops.pack_weight(w.cuda(), transposed_layout=True)
    out = torch.empty((N, 2 * H, 2 * W, Cout), device='cuda')
    xs, sk = nhwc(x), nhwc(skip)
    for py in range(2):
        for px in range(2):
            ops.conv([ops.Seg(xs)], wp, None, N, H, W, H, W, Cout, ops.taps_convT_phase(py, px), act=ACT_RELU, out=out,
                     out_place=(2 * H, 2 * W, 2, py, 2, px), res_post=sk)
    assert rel_err(nchw(out), ref) < TIGHT


@pytest.mark.parametrize('with_state', [False, True])
def test_convlstm_fp32(with_state):
    ops = _ops()
    from ess_b200._lib import EPI_LSTM
    from ess_b200.e2vid import _interleave
    g = torch.Generator().manual_seed(3)
    N, H, W, C = 2, 11, 13, 32
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(4 * C, 2 * C, 3, 3, generator=g) * 0.05
    b = torch.randn(4 * C, generator=g) * 0.1
    prev = (torch.randn(N, C, H, W, generator=g), torch.randn(N, C, H, W, generator=g)) if with_state else None
    h_ref, c_ref = O.convlstm(x, prev, {'p.Gates.weight': w, 'p.Gates.bias': b}, 'p')
    wc = w.cuda()
    if with_state:
        segs = [ops.Seg(nhwc(x)), ops.Seg(nhwc(prev[0]))]
        wp, cp = ops.pack_weight(wc, interleave=4), nhwc(prev[1])
    else:
        segs = [ops.Seg(nhwc(x))]
        wp, cp = ops.pack_weight(wc[:, :C].contiguous(), interleave=4), None
    h, c, _, _ = ops.conv(segs, wp, _interleave(b.cuda(), 4), N, H, W, H, W, 4 * C, ops.taps_conv(3, 1),
                          epilogue=EPI_LSTM, aux0=cp)
    assert rel_err(nchw(h), h_ref) < TIGHT and rel_err(nchw(c), c_ref) < TIGHT


def test_wgrad_fp32():
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    N, H, W, C0, C1, Cout = 2, 14, 18, 64, 32, 40
    a0 = torch.randn(N, C0, H // 2, W // 2, generator=g)
    a1 = torch.randn(N, C1, H, W, generator=g)
    dy = torch.randn(N, Cout, H, W, generator=g)
    w = torch.zeros(Cout, C0 + C1, 3, 3, requires_grad=True)
    b = torch.zeros(Cout, requires_grad=True)
    a = torch.cat([a0.repeat_interleave(2, 2).repeat_interleave(2, 3), a1], 1)
    out = F.conv2d(a, w, b, padding=1)
    gw, gb = torch.autograd.grad(out, [w, b], dy)
    segs = [ops.Seg(nhwc(a0), ups=1), ops.Seg(nhwc(a1))]
    dw, db = ops.wgrad(segs, nhwc(dy), N, H, W, H, W, Cout, ops.taps_conv(3, 1))
    assert rel_err(dw.view(Cout, C0 + C1, 3, 3).cpu(), gw) < TIGHT
    assert rel_err(db.cpu(), gb) < TIGHT


def test_dgrad_via_gather_conv():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    N, H, W, Cin, Cout = 2, 10, 9, 24, 32
    x = torch.randn(N, Cin, H, W, generator=g, requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1
    dy = torch.randn(N, Cout, H, W, generator=g)
    gx, = torch.autograd.grad(F.conv2d(x, w, None, padding=1), [x], dy)
    wp = ops.pack_weight(w.cuda(), swap_io=True)
    dtaps = [(-a, -b_, t) for (a, b_, t) in ops.taps_conv(3, 1)]
    dA, _, _, _ = ops.conv([ops.Seg(nhwc(dy))], wp, None, N, H, W, H, W, Cin, dtaps)
    assert rel_err(nchw(dA), gx) < TIGHT


@pytest.mark.parametrize('relu,ups,C', [(True, 0, 32), (False, 0, 256), (True, 1, 64), (True, 0, 128)])
def test_instance_norm_backward(relu, ups, C):
    ops = _ops()
    g = torch.Generator().manual_seed(6)
    N, H, W = 2, 9, 14
    y = (torch.randn(N, C, H, W, generator=g) * 1.7 + 0.3).requires_grad_(True)
    a = F.instance_norm(y, eps=1e-5)
    if relu:
        a = torch.relu(a)
    if ups:
        a = a.repeat_interleave(2, 2).repeat_interleave(2, 3)
    dA = torch.randn(a.shape, generator=g)
    gy, = torch.autograd.grad(a, [y], dA)
    yd = y.detach()
    mean = yd.mean((2, 3)).cuda().contiguous()
    rstd = (1.0 / torch.sqrt(yd.var((2, 3), unbiased=False) + 1e-5)).cuda().contiguous()
    out = ops.in_backward(nhwc(dA), nhwc(yd), mean, rstd, relu=relu, ups=ups)
    assert rel_err(nchw(out), gy) < 1e-4
    # forward materialisation
    res = torch.randn(N, C, H, W, generator=g)
    fwd = ops.norm_act_add(nhwc(yd), mean, rstd, relu=relu, res=nhwc(res))
    ref = F.instance_norm(yd, eps=1e-5)
    ref = (torch.relu(ref) if relu else ref) + res
    assert rel_err(nchw(fwd), ref) < TIGHT


@pytest.mark.parametrize('K,losses', [(11, ('dice', 'cross_entropy')), (6, ('dice',)), (19, ('cross_entropy',)),
                                      (3, ('dice', 'cross_entropy'))])
def test_task_loss_kernels(K, losses):
    ops = _ops()
    g = torch.Generator().manual_seed(7)
    N, H, W = 2, 17, 23
    logits = (torch.randn(N, K, H, W, generator=g) * 3).requires_grad_(True)
    target = torch.randint(0, K, (N, H, W), generator=g)
    target[:, :3] = 255
    ref = O.task_loss(logits, target, K, 255, losses)
    gref, = torch.autograd.grad(ref * 0.7, [logits])
    lg = nhwc(logits.detach())
    sums = ops.task_loss_sums(lg, target.cuda(), K, 255)
    loss = ops.task_loss_finish(sums, K, 255, 'dice' in losses, 'cross_entropy' in losses)
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    dl = ops.task_loss_bwd(lg, target.cuda(), K, 255, sums, 'dice' in losses, 'cross_entropy' in losses,
                           torch.tensor([0.7], device='cuda'))
    assert rel_err(nchw(dl), gref) < 1e-4
    conf = ops.confusion_logits(lg, target.cuda(), K, 255)
    assert torch.equal(conf.cpu(), O.confusion_matrix(logits.detach().argmax(1), target, K, 255))   # bit-exact
    conf2 = ops.confusion_labels(logits.detach().argmax(1).cuda(), target.cuda(), K, 255)
    assert torch.equal(conf2.cpu(), conf.cpu())


def test_task_loss_all_ignored_and_module():
    import ess_b200
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=4, ignore_index=255)
    logits = torch.randn(1, 4, 6, 6, device='cuda', requires_grad=True)
    target = torch.randint(0, 4, (1, 6, 6), device='cuda')
    loss = crit(logits, target)
    loss.backward()
    ref_logits = logits.detach().cpu().requires_grad_(True)
    ref = O.task_loss(ref_logits, target.cpu(), 4, 255)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5
    assert rel_err(logits.grad, ref_logits.grad) < 1e-4


@pytest.mark.parametrize('H,W,C', [(16, 24, 5), (30, 43, 3), (200, 346, 5)])
def test_event_prepare(H, W, C):
    ops = _ops()
    from ess_b200.reconstructor import crop_padding
    g = torch.Generator().manual_seed(8)
    B, T = 2, 3
    data = torch.randn(B, T * C, H, W, generator=g) * (torch.rand(B, T * C, H, W, generator=g) < 0.2)
    dd = data.cuda()
    stats = ops.event_stats(dd, T, C)
    left, right, top, bottom = crop_padding(H, W, 3)
    assert (left, right, top, bottom) == O.crop_padding(H, W, 3)
    for i in range(T):
        win = data[:, i * C:(i + 1) * C]
        ref = O.reflect_pad(O.event_normalize(win), 3)
        out = ops.event_prepare(dd[:, i * C:(i + 1) * C], stats[i], True, H + top + bottom, W + left + right, top, left,
                                8)
        assert rel_err(nchw(out)[:, :C], ref) < 1e-5
        assert float(out[..., C:].abs().max()) == 0.0
    # all-zero window: passes through unchanged (inference_utils.py:100)
    z = torch.zeros(1, C, H, W, device='cuda')
    st = ops.event_stats(z, 1, C)
    out = ops.event_prepare(z, st[0], True, H + top + bottom, W + left + right, top, left, 8)
    assert float(out.abs().max()) == 0.0


def test_layout_and_bilinear():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 5, 7, 9, generator=g)
    out = ops.nchw_to_nhwc(x.cuda(), 8)
    assert torch.equal(out[..., :5].cpu(), x.permute(0, 2, 3, 1)) and float(out[..., 5:].abs().max()) == 0
    y = torch.randn(2, 6, 5, 7, generator=g)
    up = ops.bilinear_up2(nhwc(y))
    ref = F.interpolate(y, scale_factor=2, mode='bilinear', align_corners=False)
    assert rel_err(nchw(up), ref) < 1e-6
    ch = torch.randn(2, 8, 6, 4, generator=g)
    s = ops.upsample2_bwd(nhwc(ch), 3, 2, 8)
    assert rel_err(nchw(s), F.avg_pool2d(ch, 2) * 4) < 1e-6


def test_radam_kernel_vs_oracle():
    import ess_b200.optim
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(33, 17, device='cuda'))
    w2 = w.detach().cpu().clone()
    opt = ess_b200.optim.RAdam([w], lr=5e-4, weight_decay=0., betas=(0., 0.999))
    state = {}
    for i in range(9):       # crosses the N_sma >= 5 rectification switch (radam.py:60-66)
        g = torch.randn(33, 17)
        w.grad = g.cuda()
        opt.step()
        O.radam_step(w2, g, state, 5e-4, (0., 0.999))
        assert rel_err(w.detach(), w2) < 1e-6, i


def test_radam_multi_tensor_exponential_lr_and_state_roundtrip():
    """The multi-tensor launch over a realistic parameter set (60 tensors of mixed sizes incl. sizes that are not
    multiples of 4 -> two launches, vector and scalar paths), with gradient views into ONE flat bucket (ess_b200.dp)
    whose slices are not 16 B aligned, driven through torch's ExponentialLR exactly as the reference does
    (training/base_trainer.py:64-66: one scheduler step per epoch) and a state_dict save / load in the middle
    (utils/saver.py) -- against the oracle's per-tensor restatement of utils/radam.py:15-80."""
    import ess_b200.optim
    g = torch.Generator().manual_seed(1)
    shapes = [(256, 256, 3, 3), (256,), (11, 32, 1, 1), (11,), (7, 5), (3,), (130,)] + [(64, 9), (33,)] * 26 + [(1,)]
    assert len(shapes) == 60
    params = [torch.nn.Parameter((torch.randn(s, generator=g) * 0.1).cuda()) for s in shapes]
    ref = [p.detach().cpu().clone() for p in params]
    states = [{} for _ in params]
    total = sum(p.numel() for p in params)
    flat = torch.zeros(total, device='cuda')
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    opt = ess_b200.optim.RAdam(params, lr=5e-4, weight_decay=1e-2, betas=(0., 0.999))
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)
    lr = 5e-4
    from ess_b200 import _lib
    for epoch in range(3):
        for it in range(4):
            gs = [torch.randn(s, generator=g) for s in shapes]
            flat.copy_(torch.cat([x.flatten() for x in gs]).cuda())
            l0 = _lib.launch_count
            opt.step()
            assert _lib.launch_count - l0 == 2                      # 60 tensors -> 48 + 12
            for r, x, st in zip(ref, gs, states):
                O.radam_step(r, x, st, lr, (0., 0.999), weight_decay=1e-2)
        sched.step()
        lr *= 0.5
        assert abs(opt.param_groups[0]['lr'] - lr) < 1e-12
        if epoch == 0:      # checkpoint round trip: a fresh optimizer continues from the saved state
            sd = opt.state_dict()
            opt = ess_b200.optim.RAdam(params, lr=5e-4, weight_decay=1e-2, betas=(0., 0.999))
            opt.load_state_dict(sd)
            sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5, last_epoch=0)
            assert abs(opt.param_groups[0]['lr'] - lr) < 1e-12
    for p, r in zip(params, ref):
        assert rel_err(p.detach(), r) < 1e-5, tuple(p.shape)
    assert all(p._version > 0 for p in params)                      # packed-weight caches see the update


def test_voxel_grids():
    from ess_b200.voxel import VoxelGrid, generate_voxel_grid
    g = torch.Generator().manual_seed(0)
    n, C, H, W = 20000, 5, 40, 64
    x = torch.rand(n, generator=g) * (W + 1) - 0.7
    y = torch.rand(n, generator=g) * (H + 1) - 0.7
    pol = (torch.rand(n, generator=g) > 0.5).float()
    t = torch.sort(torch.rand(n, generator=g))[0] * 1e3
    ref = O.voxel_grid_dsec(x, y, pol, t, C, H, W)
    out = VoxelGrid(C, H, W, False).convert(x.cuda(), y.cuda(), pol.cuda(), t.cuda())
    assert float((out.cpu() - ref).abs().max()) < 1e-4          # float atomics: summation-order noise only
    ev = torch.stack([torch.floor(torch.rand(n, generator=g) * (W + 2)) - 1, torch.floor(torch.rand(n, generator=g) * (H + 2)) - 1,
                      torch.sort(torch.rand(n, generator=g))[0] * 50.0, (torch.rand(n, generator=g) > 0.5).float()], 1).double()
    for sep in (True, False):
        ref = O.voxel_grid_ddd17(ev.numpy(), (H, W), C, separate_pol=sep)
        out = generate_voxel_grid(ev.cuda(), (H, W), C, separate_pol=sep)
        assert out.shape == ref.shape and float((out.cpu() - ref).abs().max()) < 1e-4
    # degenerate: all events share one timestamp (deltaT == 0 branch, data_util.py:75-76)
    ev2 = ev.clone()
    ev2[:, 2] = 7.0
    assert float((generate_voxel_grid(ev2.cuda(), (H, W), C).cpu() - O.voxel_grid_ddd17(ev2.numpy(), (H, W), C)).abs().max()) < 1e-4


# ------------------------------------------------------------------------------ 1x1 streaming kernels
@pytest.mark.parametrize('N,H,W,Cin,K,norm', [
    (2, 9, 7, 32, 11, True),          # classifier shape, ragged block (126 rows)
    (1, 40, 56, 32, 6, True),         # DDD17 class count, several blocks
    (3, 17, 33, 64, 16, False),       # Cin = 64, Cout = 16, rows cross sample boundaries inside a warp
    (2, 16, 24, 32, 1, False),        # E2VID prediction layer (Cout = 1)
])
def test_pw_conv_fwd_dgrad_wgrad(N, H, W, Cin, K, norm):
    """essb_pw_conv_{fwd,dgrad,wgrad} vs autograd of conv1x1(relu(instance_norm(y))) (style_networks.py:34,88)."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    y = (torch.randn(N, Cin, H, W, generator=g) * 1.5 + 0.3).double()
    w = (torch.randn(K, Cin, 1, 1, generator=g) * 0.2).double().requires_grad_(True)
    b = torch.randn(K, generator=g).double().requires_grad_(True)
    gy = torch.randn(N, K, H, W, generator=g).double()
    a = (torch.relu(F.instance_norm(y, eps=1e-5)) if norm else y).detach().requires_grad_(True)
    out = F.conv2d(a, w, b)
    out.backward(gy)
    yd = nhwc(y.float())
    mean = rstd = None
    if norm:
        mean, rstd = ops.in_stats(yd)
    seg = ops.Seg(yd, mean=mean, rstd=rstd, relu=norm)
    wd = w.detach().float().reshape(K, Cin).cuda().contiguous()
    o = ops.pw_conv_fwd(seg, wd, b.detach().float().cuda(), N, H, W, K)
    assert rel_err(nchw(o), out) < TIGHT
    gyd = nhwc(gy.float())
    dx = ops.pw_conv_dgrad(gyd, wd, Cin)
    assert rel_err(nchw(dx), a.grad) < TIGHT
    dw, db = ops.pw_conv_wgrad(seg, gyd)
    assert rel_err(dw.cpu(), w.grad.reshape(K, Cin)) < TIGHT
    assert rel_err(db.cpu(), b.grad) < TIGHT


def test_pw_conv_sigmoid_matches_pred_layer():
    ops = _ops()
    from ess_b200._lib import ACT_SIGMOID
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 32, 24, 40, generator=g)
    w = torch.randn(1, 32, 1, 1, generator=g) * 0.3
    b = torch.randn(1, generator=g)
    ref = torch.sigmoid(F.conv2d(x, w, b))
    o = ops.pw_conv_fwd(ops.Seg(nhwc(x)), w.reshape(1, 32).cuda().contiguous(), b.cuda(), 2, 24, 40, 1, act=ACT_SIGMOID)
    assert rel_err(nchw(o), ref) < TIGHT


@pytest.mark.parametrize('B,T,count', [(2, 3, 5 * 16 * 24), (3, 2, 1001), (1, 4, 70000)])
def test_event_stats_vectorised(B, T, count):
    """essb_event_stats (128-bit path, scalar tail path, multi-chunk slabs) vs torch sums."""
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, T, count, generator=g) * (torch.rand(B, T, count, generator=g) < 0.3)
    xd = x.cuda()
    from ess_b200.ops import _p, _stream, call
    stats = torch.empty((T, 3), device='cuda', dtype=torch.float64)
    call('essb_event_stats', _p(xd), xd.stride(0), B, T, count, _p(stats), _stream())
    xd64 = x.double()
    ref = torch.stack([xd64.sum((0, 2)), (xd64 * xd64).sum((0, 2)), (x != 0).double().sum((0, 2))], 1)
    assert rel_err(stats.cpu(), ref) < 1e-6
    assert torch.equal(stats[:, 2].cpu(), ref[:, 2])


def test_split_bf16_zero_pads_channels():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 6, 10, 32, generator=g).cuda()
    hi = torch.full((2, 6, 10, 64), 7.0, device='cuda', dtype=torch.bfloat16)
    lo = torch.full_like(hi, 7.0)
    ops.split_bf16(ops.Seg(x), 2, 6, 10, hi, lo, 0, c_pad=64)
    assert float(hi[..., 32:].abs().max()) == 0.0 and float(lo[..., 32:].abs().max()) == 0.0
    assert rel_err(hi[..., :32].float() + lo[..., :32].float(), x) < 1e-4


@pytest.mark.parametrize('N,H,W', [(2, 32, 64), (1, 46, 70), (3, 16, 24)])
def test_stem_conv_fwd_wgrad(N, H, W):
    """essb_stem_conv_{fwd,wgrad} (ResNet-18 conv1: 1 -> 64, 7x7, stride 2, pad 3; style_networks.py:117-121) vs
    autograd, incl. ragged tiles (OW = 35 not a multiple of the 32-pixel tile, OH = 23 not a multiple of 8)."""
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    x = torch.rand(N, 1, H, W, generator=g).double()
    w = (torch.randn(64, 1, 7, 7, generator=g) * 0.1).double().requires_grad_(True)
    y = F.conv2d(x, w, stride=2, padding=3)
    gy = torch.randn(y.shape, generator=g).double()
    y.backward(gy)
    assert ops.stem_conv_supported(1, 64, 7, 2)
    xd = nhwc(x.float())
    out = ops.stem_conv_fwd(xd, w.detach().float().cuda().contiguous(), 2, 3)
    assert rel_err(nchw(out), y) < TIGHT
    dw = ops.stem_conv_wgrad(xd, nhwc(gy.float()), 64, 7, 2, 3)
    assert rel_err(dw.cpu(), w.grad) < TIGHT
