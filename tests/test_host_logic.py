"""Host-side logic of the drop-in modules, checked on the CPU (no kernel calls): the tap lists that turn strided /
transposed convolutions and their gradients into gather-convolutions, the reflect-pad geometry, the spatial tile
choice, the RAdam step-size schedule and the data-parallel sharding.  Each gather list is *executed* by a tiny
pure-PyTorch emulation of the kernels' gather semantics and compared with torch's own operator."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import O
from ess_b200 import dp, ops
from ess_b200.optim import rectification
from ess_b200.reconstructor import crop_padding


def gather_conv(x, w_taps, taps, OH, OW, stride=1):
    """out[n, co, oy, ox] = sum_t sum_ci x[n, ci, oy*stride + dy_t, ox*stride + dx_t] * w_taps[widx_t][co, ci]
    (zero outside the image) -- the semantics of essb_conv_fp32 / essb_conv_tc_run."""
    N, Cin, H, W = x.shape
    out = torch.zeros(N, w_taps[0].shape[0], OH, OW, dtype=x.dtype)
    pad = 8
    xp = F.pad(x, (pad, pad, pad, pad))
    for (dy, dx, wi) in taps:
        ys = torch.arange(OH) * stride + dy + pad
        xs = torch.arange(OW) * stride + dx + pad
        ok_y = (ys >= 0) & (ys < H + 2 * pad)
        ok_x = (xs >= 0) & (xs < W + 2 * pad)
        patch = xp[:, :, ys.clamp(0, H + 2 * pad - 1)][:, :, :, xs.clamp(0, W + 2 * pad - 1)]
        patch = patch * ok_y.view(1, 1, -1, 1) * ok_x.view(1, 1, 1, -1)
        out += torch.einsum('nchw,oc->nohw', patch, w_taps[wi])
    return out


@pytest.mark.parametrize('k,pad,stride', [(3, 1, 1), (5, 2, 1), (5, 2, 2), (1, 0, 1), (7, 3, 2)])
def test_taps_conv_reproduce_conv2d(k, pad, stride):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 12, 14, generator=g, dtype=torch.float64)
    w = torch.randn(4, 3, k, k, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=stride, padding=pad)
    w_taps = [w[:, :, t // k, t % k] for t in range(k * k)]
    out = gather_conv(x, w_taps, ops.taps_conv(k, pad), ref.shape[2], ref.shape[3], stride)
    assert torch.allclose(out, ref, atol=1e-12)


def test_transposed_conv_phases_reproduce_conv_transpose2d():
    """ConvTranspose2d(k=5, stride 2, padding 2, output_padding 1) (e2vid/model/submodules.py:39) as four
    sub-pixel phases of ordinary gather-convolutions (ops.taps_convT_phase)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 7, 9, generator=g, dtype=torch.float64)
    w = torch.randn(3, 4, 5, 5, generator=g, dtype=torch.float64)       # [Cin, Cout, k, k]
    ref = F.conv_transpose2d(x, w, stride=2, padding=2, output_padding=1)
    w_taps = [w[:, :, t // 5, t % 5].t() for t in range(25)]            # [Cout, Cin] per tap
    out = torch.zeros_like(ref)
    n_taps = 0
    for py in range(2):
        for px in range(2):
            taps = ops.taps_convT_phase(py, px)
            n_taps += len(taps)
            out[:, :, py::2, px::2] = gather_conv(x, w_taps, taps, 7, 9)
    assert n_taps == 25 and torch.allclose(out, ref, atol=1e-12)


def test_merged_transposed_conv_phases_reproduce_conv_transpose2d():
    """ops.merge_convT_phases: ONE 3x3 gather-convolution with 4*Cout output channels whose channel block j is sub-pixel
    phase (j >> 1, j & 1) == ConvTranspose2d(k5, s2, p2, output_padding 1) (e2vid/model/submodules.py:39-40); the
    placement rule is the one the tcgen05 epilogue applies (essb_conv_tc.phase_cout)."""
    g = torch.Generator().manual_seed(5)
    N, Cin, Cout, H, W = 2, 3, 4, 5, 6
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
    wt = torch.randn(Cin, Cout, 5, 5, generator=g).double()        # fp32-representable: the merge keeps fp32 weights
    ref = F.conv_transpose2d(x, wt, stride=2, padding=2, output_padding=1)
    wm = ops.merge_convT_phases(wt.float()).double()            # [4*Cout, Cin, 3, 3]
    assert wm.shape == (4 * Cout, Cin, 3, 3) and int((wm.abs().sum((1, 2, 3)) == 0).sum()) == 0
    taps = ops.taps_conv(3, 1)
    y = gather_conv(x, [wm[:, :, t // 3, t % 3] for t in range(9)], taps, H, W)
    out = torch.zeros_like(ref)
    for j in range(4):
        out[:, :, (j >> 1)::2, (j & 1)::2] = y[:, j * Cout:(j + 1) * Cout]
    assert float((out - ref).abs().max()) < 1e-12
    used = sum(len(ops.taps_convT_phase(py, px)) for py in range(2) for px in range(2))
    assert used == 25 and int((wm.abs().sum((0, 1)) > 0).sum()) == 9   # 25 of 36 (phase, offset) blocks carry weights


@pytest.mark.parametrize('N,H,W,expect', [(8, 440, 640, [(220, False), (112, False), (56, True)]),
                                          (8, 200, 352, [(104, False), (52, True), (26, True)]),
                                          (1, 440, 640, [(220, False), (110, False), (55, False)]),
                                          (2, 64, 96, [(32, False), (16, False), (8, False)])])
def test_row_stack_plan(N, H, W, expect, monkeypatch):
    """E2VIDRecurrent._stack_plan: a level is row-stacked (rows per image > its height) only when that saves 16-row output
    patches for its ConvLSTM or lets the next level's stride-2 conv run on the tall view (input period = 2 x output
    period); never for a single image; ESS_B200_STACK=0 switches it off."""
    from ess_b200.e2vid import E2VIDRecurrent as E
    monkeypatch.delenv('ESS_B200_STACK', raising=False)
    plan = E._stack_plan(N, H, W, 3)
    assert plan == expect
    for i, (rows, tall) in enumerate(plan):
        oh = H >> (i + 1)
        assert rows >= oh
        if tall:
            assert i > 0 and plan[i - 1][0] == 2 * rows and rows > oh
        if rows > oh:     # never more patches than the dense layout
            assert -(-N * rows // 16) <= N * -(-oh // 16)
    monkeypatch.setenv('ESS_B200_STACK', '0')
    assert E._stack_plan(N, H, W, 3) == [(H >> (i + 1), False) for i in range(3)]


@pytest.mark.parametrize('k,pad,stride', [(3, 1, 2), (1, 0, 2), (5, 2, 2), (3, 1, 1)])
def test_dgrad_phase_taps_reproduce_the_input_gradient(k, pad, stride):
    g = torch.Generator().manual_seed(2)
    H, W = 12, 16
    x = torch.randn(2, 3, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(4, 3, k, k, generator=g, dtype=torch.float64)
    y = F.conv2d(x, w, stride=stride, padding=pad)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    gx_ref, = torch.autograd.grad(y, x, gy)
    w_taps = [w[:, :, t // k, t % k].t() for t in range(k * k)]          # dgrad: swap in/out channels
    gx = torch.zeros_like(gx_ref)
    for (py, px), taps in ops.dgrad_phase_taps(k, pad, stride).items():
        if taps:
            gx[:, :, py::stride, px::stride] = gather_conv(gy, w_taps, taps, H // stride, W // stride)
    assert torch.allclose(gx, gx_ref, atol=1e-12)


def test_stride2_wgrad_parity_decomposition():
    """essb_wgrad_tc(a_stride=2): input offset o = ky - pad splits into parity o & 1 and in-plane shift (o - parity) / 2
    (floor semantics for negative offsets) -- check it addresses the same input pixel."""
    for o in range(-3, 4):
        par = o & 1
        shift = (o - par) // 2
        for y in range(5):
            assert 2 * (y + shift) + par == 2 * y + o


@pytest.mark.parametrize('H,W', [(440, 640), (200, 346), (260, 346), (63, 65), (8, 8), (120, 216)])
def test_crop_padding_matches_the_oracle(H, W):
    assert crop_padding(H, W, 3) == O.crop_padding(H, W, 3)
    l, r, t, b = crop_padding(H, W, 3)
    assert (H + t + b) % 8 == 0 and (W + l + r) % 8 == 0 and 0 <= l - r <= 1 and 0 <= t - b <= 1


@pytest.mark.parametrize('OW,OH', [(320, 220), (160, 110), (80, 55), (44, 25), (640, 440), (8, 8), (1, 300)])
def test_pick_bw_log2_minimises_padding(OW, OH):
    b = ops.pick_bw_log2(OW, OH)
    def waste(bl):
        bw, bh = 1 << bl, 128 >> bl
        return math.ceil(OW / bw) * bw * math.ceil(OH / bh) * bh
    assert 2 <= b <= 6 and waste(b) == min(waste(x) for x in range(2, 7))


def test_radam_rectification_schedule_matches_the_oracle():
    """step size of ess_b200.optim.RAdam == the reference rule restated in oracle.radam_step (radam.py:53-66)."""
    for betas in ((0.0, 0.999), (0.9, 0.999)):
        p, st = torch.zeros(1, dtype=torch.float64), {}
        m = torch.zeros(1, dtype=torch.float64)
        v = torch.zeros(1, dtype=torch.float64)
        q = torch.zeros(1, dtype=torch.float64)
        for step in range(1, 40):
            grad = torch.tensor([math.sin(step) + 1.5], dtype=torch.float64)
            O.radam_step(p, grad, st, lr=1e-2, betas=betas)
            v = betas[1] * v + (1 - betas[1]) * grad * grad
            m = betas[0] * m + (1 - betas[0]) * grad
            rect, size = rectification(step, float(betas[0]), float(betas[1]))
            q = q - 1e-2 * size * (m / (v.sqrt() + 1e-8) if rect else m)
            assert abs(float(p) - float(q)) < 1e-12 * max(1.0, abs(float(p)))
        assert rectification(1, 0.0, 0.999)[0] is False and rectification(39, 0.0, 0.999)[0] is True


def test_shard_batch_partitions_the_global_batch():
    for world in (1, 2, 4, 8):
        spans = [dp.shard_batch(64, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 64
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        dp.shard_batch(10, 0, 4)


@pytest.mark.parametrize('skip_connect', [True, False])
def test_decoder_program_equals_oracle(skip_connect):
    """The static node program of ess_b200.SemSegE2VID, interpreted with plain torch ops (helpers.program_forward,
    the teacher of tests/test_gpu_teacher.py), reproduces the reference restatement O.semseg_forward -- outputs and
    every parameter gradient -- in fp64: the wiring of the executor's graph is pinned on CPU."""
    from helpers import O, make_labels, make_latents, make_semseg, program_forward
    K, B, H, W = 5, 1, 16, 24
    kw = dict(skip_connect=True, skip_type='concat') if skip_connect else dict(skip_connect=False, skip_type='sum')
    dec = make_semseg(K, input_c=64, **kw)
    lat = {k: v.double() for k, v in make_latents(B, H, W, base=8).items()}
    labels = make_labels(B, H, W, K)
    sd = {k: v.detach().double().requires_grad_(True) for k, v in dec.state_dict().items()}
    outs, _ = program_forward(dec, sd, lat)
    ref = O.semseg_forward(sd, lat, **kw)
    for k in (4, 2, 1):
        assert torch.equal(outs[k], ref[k]) or float((outs[k] - ref[k]).abs().max()) < 1e-12
    g_a = torch.autograd.grad(O.task_loss(outs[1], labels, K), list(sd.values()), allow_unused=True)
    g_b = torch.autograd.grad(O.task_loss(ref[1], labels, K), list(sd.values()), allow_unused=True)
    for n, a, b in zip(sd, g_a, g_b):
        assert (a is None) == (b is None), n
        if a is not None:
            assert float((a - b).abs().max()) <= 1e-12 * max(1.0, float(b.abs().max())), n


def test_style_encoder_pretrained_weights_policy(monkeypatch, tmp_path):
    """The reference builds its image encoder from resnet18(pretrained=True) (models/style_networks.py:117-121).
    The drop-in takes the ImageNet weights from the hub cache when they are there, and otherwise (no network in this
    container) warns LOUDLY instead of silently starting a UDA run from random init."""
    import warnings
    import torchvision.models as tvm
    import ess_b200
    from ess_b200 import style_encoder as se
    monkeypatch.setenv('ESS_B200_PRETRAINED', '1')
    monkeypatch.setattr(torch.hub, 'get_dir', lambda: str(tmp_path))
    monkeypatch.setattr(se, '_warned_pretrained', [False])

    def no_network(*a, **k):
        raise OSError('no network')
    monkeypatch.setattr(torch.hub, 'load_state_dict_from_url', no_network)
    monkeypatch.setattr(tvm._api.WeightsEnum, 'get_state_dict', lambda self, *a, **k: no_network())
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        m = ess_b200.StyleEncoderE2VID(1, skip_connect=True)
    assert m.pretrained_source == 'random-init'
    assert any('RANDOM init' in str(w.message) and issubclass(w.category, RuntimeWarning) for w in rec)
    # cache hit: a resnet18 state_dict under <hub>/checkpoints/ is loaded into the holders (bn1, layer1-3)
    r = tvm.resnet18(weights=None)
    with torch.no_grad():
        r.layer2[0].conv1.weight.fill_(0.125)
    os_dir = tmp_path / 'checkpoints'
    os_dir.mkdir()
    torch.save(r.state_dict(), str(os_dir / 'resnet18-f37072fd.pth'))
    m = ess_b200.StyleEncoderE2VID(1, skip_connect=True)
    assert m.pretrained_source == 'cache'
    assert float(m.encoder_scale_2[0].conv1.weight.min()) == 0.125 == float(m.encoder_scale_2[0].conv1.weight.max())
    monkeypatch.setenv('ESS_B200_PRETRAINED', '0')
    assert ess_b200.StyleEncoderE2VID(1).pretrained_source == 'disabled'


def test_no_undefined_names_in_python_sources():
    """The CUDA paths cannot execute in the CPU-only build container; tools/lint.py at least proves that no Python
    source loads a name that is never bound (the round-2 `sink` NameError reached the GPU box once)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'lint.py')], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
