"""Thin tensor-level wrappers over the C ABI (libess_b200.so).

Activations are "pixel-major" NHWC fp32 tensors of shape [N, H, W, C] (contiguous).  PyTorch is used
only to own device memory and streams; every arithmetic op on the hot path is a call into the
library.  Nothing here falls back to torch ops: a missing library or a failing call raises.
"""
import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_SIGMOID, EPI_GRU_OUT, EPI_GRU_UR, EPI_LINEAR, EPI_LSTM, Conv, ConvTc,
                   Src, TcView, Wgrad, call)

IN_EPS = 1e-5


def _stream(device=None):
    """The launching stream: the CURRENT device's current stream.  Every module entry point runs under
    `on_device_of(operand)`, and `require_cuda` refuses operands of another device, so the current device is the
    operands' device and kernels never launch with foreign pointers."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device_of(t):
    """Context manager making `t`'s CUDA device current (the reference places its modules with
    `settings.gpu_device = 'cuda:1'` and never calls torch.cuda.set_device).  Raises for CPU tensors: no fallback."""
    require_cuda_any(t)
    return torch.cuda.device(t.device)


def require_cuda_any(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('ess_b200 runs on CUDA (sm_100a) only; got a %s tensor -- there is no CPU '
                               'fallback' % t.device)


def _p(t, offset_elems=0):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def require_cuda(*ts):
    """Operands must be CUDA tensors of the CURRENT device (see _stream / on_device_of)."""
    require_cuda_any(*ts)
    cur = None
    for t in ts:
        if t is None:
            continue
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError('ess_b200: operand on %s but the current CUDA device is cuda:%d (module entry points '
                               'switch to their operands\' device; mixed-device operands are not supported)' % (t.device, cur))


@dataclass
class Seg:
    """One input channel segment of a convolution (see essb_src in include/ess_b200.h)."""
    t: torch.Tensor                 # [N, Hs, Ws, ld] fp32 NHWC (Hs = H >> ups)
    C: Optional[int] = None         # channels used (default: all)
    c_off: int = 0                  # first channel inside t
    ups: int = 0
    mean: Optional[torch.Tensor] = None   # [N, C]
    rstd: Optional[torch.Tensor] = None
    relu: bool = False

    def fill(self, s: Src):
        c = self.C if self.C is not None else self.t.shape[-1] - self.c_off
        s.ptr = _p(self.t, self.c_off)
        s.mean = _p(self.mean)
        s.rstd = _p(self.rstd)
        s.ld = self.t.shape[-1]
        s.C = c
        s.ups = self.ups
        s.relu = 1 if self.relu else 0
        return c


def taps_conv(k, pad):
    """(dy, dx, widx) of a k x k convolution with zero padding `pad` (weights in [.., ky, kx] order)."""
    return [(ky - pad, kx - pad, ky * k + kx) for ky in range(k) for kx in range(k)]


def taps_convT_phase(py, px, k=5, pad=2):
    """Taps of output phase (py, px) of ConvTranspose2d(k, stride 2, pad, output_padding 1):
    oy = 2*iy - pad + ky  =>  for oy = 2*oy' + py: ky = py + pad (mod 2), iy = oy' + (py + pad - ky)/2."""
    out = []
    for ky in range(k):
        if (py + pad - ky) % 2:
            continue
        for kx in range(k):
            if (px + pad - kx) % 2:
                continue
            out.append(((py + pad - ky) // 2, (px + pad - kx) // 2, ky * k + kx))
    return out


def _fill_taps(d, taps, with_widx=True):
    d.ntaps = len(taps)
    for i, tp in enumerate(taps):
        d.dy[i], d.dx[i] = tp[0], tp[1]
        if with_widx:
            d.widx[i] = tp[2]


def pack_weight(w, scale=None, transposed_layout=False, swap_io=False, flip=False, interleave=1):
    """Reference-layout conv weight -> packed fp32 [T][Kin][NoutP] (essb_pack_weight)."""
    require_cuda(w)
    w = w.detach().contiguous().float()
    if transposed_layout:
        cin, cout = w.shape[0], w.shape[1]
    else:
        cout, cin = w.shape[0], w.shape[1]
    T = w.shape[2] * w.shape[3]
    kin, nout = (cout, cin) if swap_io else (cin, cout)
    out = torch.empty((T, kin, (nout + 3) // 4 * 4), device=w.device, dtype=torch.float32)
    call('essb_pack_weight', _p(w), _p(scale), _p(out), cout, cin, T, int(transposed_layout), int(swap_io),
         int(flip), int(interleave), _stream())
    return out


def conv(segs: Sequence[Seg], w_packed, bias, N, H, W, OH, OW, Cout, taps, stride=1, epilogue=EPI_LINEAR,
         act=ACT_NONE, out=None, out2=None, res_pre=None, res_post=None, aux0=None, aux1=None, want_stats=False,
         out_place: Optional[Tuple[int, int, int, int, int, int]] = None, accumulate=False, planes=None):
    """essb_conv_fp32.  Returns (out, out2, stats_partial, tiles)."""
    d = Conv()
    segs[0].fill(d.src[0])
    if len(segs) > 1:
        segs[1].fill(d.src[1])
    dev = segs[0].t.device
    d.w, d.bias = _p(w_packed), _p(bias)
    d.N, d.H, d.W, d.OH, d.OW, d.Cout = N, H, W, OH, OW, Cout
    d.sy = d.sx = stride
    if out_place is None:
        d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = OH, OW, 1, 0, 1, 0
    else:
        d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = out_place
    if epilogue == EPI_LINEAR:
        if out is None:
            out = torch.empty((N, d.OHf, d.OWf, Cout), device=dev, dtype=torch.float32)
        d.ldo = out.shape[-1]
    else:
        hidden = {EPI_LSTM: Cout // 4, EPI_GRU_UR: Cout // 2, EPI_GRU_OUT: Cout}[epilogue]
        if out is None:
            out = torch.empty((N, OH, OW, hidden), device=dev, dtype=torch.float32)
        if out2 is None and epilogue in (EPI_LSTM, EPI_GRU_UR):
            out2 = torch.empty((N, OH, OW, hidden), device=dev, dtype=torch.float32)
        d.ldo = hidden
    d.out, d.out2 = _p(out), _p(out2)
    d.res_pre, d.res_post = _p(res_pre), _p(res_post)
    r = res_pre if res_pre is not None else res_post
    d.ld_res = r.shape[-1] if r is not None else 0
    d.aux0, d.aux1 = _p(aux0), _p(aux1)
    d.accumulate = int(accumulate)
    d.epilogue, d.act = epilogue, act
    _fill_taps(d, taps)
    stats = None
    tiles = _lib.lib().essb_conv_tiles_per_sample(C.byref(d))
    if want_stats:
        stats = torch.empty((N, tiles, Cout, 2), device=dev, dtype=torch.float32)
        d.stats_partial = _p(stats)
    if planes is not None:
        d.out_hi, d.out_lo, d.ld_planes = _p(planes[0]), _p(planes[1]), planes[0].shape[-1]
    call('essb_conv_fp32', C.byref(d), _stream())
    return out, out2, stats, tiles


def in_finalize(stats_partial, count):
    N, tiles, Cc, _ = stats_partial.shape
    mean = torch.empty((N, Cc), device=stats_partial.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    call('essb_in_finalize', _p(stats_partial), N, tiles, Cc, count, IN_EPS, _p(mean), _p(rstd), _stream())
    return mean, rstd


def norm_act_add(y, mean=None, rstd=None, relu=False, res=None, out=None, c_off=0, C_=None):
    """out = act((y - mean) * rstd) + res over [N, H, W, C]."""
    N, H, W = y.shape[:3]
    Cc = C_ if C_ is not None else y.shape[-1] - c_off
    if out is None:
        out = torch.empty((N, H, W, Cc), device=y.device, dtype=torch.float32)
    call('essb_norm_act_add', _p(y, c_off), y.shape[-1], _p(mean), _p(rstd), int(relu), _p(res),
         res.shape[-1] if res is not None else 0, _p(out), out.shape[-1], N, H * W, Cc, _stream())
    return out


def in_backward(dA, y, mean, rstd, relu, ups=0, c_off=0, extra=None, planes_ld=0, want_fp32=True):
    """Gradient w.r.t. y of  A = [upsample2]( relu?( (y-mean)*rstd ) )  given dA (channel slice
    [c_off, c_off+C) of dA's last dim).  y: [N, H, W, C].  planes_ld > 0: the second pass also (or, with
    want_fp32=False, only) writes the gradient as bf16 hi/lo planes of pitch planes_ld (zero padded) and a
    (dy | None, (hi, lo)) pair is returned."""
    N, H, W, Cc = y.shape
    dev = y.device
    blocks = _lib.lib().essb_in_bwd_blocks(H * W)
    g = torch.empty((N, H, W, Cc), device=dev, dtype=torch.float32)
    partial = torch.empty((N, blocks, Cc, 2), device=dev, dtype=torch.float32)
    call('essb_in_bwd_pass1', _p(dA, c_off), dA.shape[-1], ups, _p(extra), extra.shape[-1] if extra is not None else 0,
         _p(y), Cc, _p(mean), _p(rstd), int(relu), _p(g), _p(partial), N, H, W, Cc, _stream())
    totals = torch.empty((N, Cc, 2), device=dev, dtype=torch.float32)
    call('essb_partial_reduce', _p(partial), N, blocks, Cc, _p(totals), _stream())
    dy = torch.empty_like(g) if (want_fp32 or not planes_ld) else None
    planes = None
    if planes_ld:
        planes = (torch.empty((N, H, W, planes_ld), device=dev, dtype=torch.bfloat16),
                  torch.empty((N, H, W, planes_ld), device=dev, dtype=torch.bfloat16))
    call('essb_in_bwd_pass2', _p(g), _p(y), Cc, _p(mean), _p(rstd), _p(totals), _p(dy),
         _p(planes[0]) if planes else None, _p(planes[1]) if planes else None, planes_ld, N, H * W, Cc, _stream())
    return (dy, planes) if planes_ld else dy


def upsample2_bwd(dA, H, W, Cc, c_off=0, out=None, accumulate=False):
    """Sum of the 2x2 children: dA [N, 2H, 2W, ld] (channel slice) -> [N, H, W, C]."""
    N = dA.shape[0]
    if out is None:
        out = torch.empty((N, H, W, Cc), device=dA.device, dtype=torch.float32)
        accumulate = False
    call('essb_upsample2_bwd', _p(dA, c_off), dA.shape[-1], _p(out), out.shape[-1], N, H, W, Cc, int(accumulate),
         _stream())
    return out


def wgrad(segs: Sequence[Seg], dy, N, H, W, OH, OW, Cout, taps, stride=1, want_bias=True):
    """essb_wgrad_fp32 -> (dW [Cout, Cin_total, ntaps], dbias [Cout] | None)."""
    d = Wgrad()
    cin = segs[0].fill(d.src[0])
    if len(segs) > 1:
        cin += segs[1].fill(d.src[1])
    dev = dy.device
    d.dy_ptr = _p(dy)
    d.N, d.H, d.W, d.OH, d.OW, d.Cout, d.ld_dy = N, H, W, OH, OW, Cout, dy.shape[-1]
    d.sy = d.sx = stride
    _fill_taps(d, taps, with_widx=False)
    dw = torch.empty((Cout, cin, len(taps)), device=dev, dtype=torch.float32)
    db = torch.empty((Cout,), device=dev, dtype=torch.float32) if want_bias else None
    d.dw, d.dbias = _p(dw), _p(db)
    nbytes = _lib.lib().essb_wgrad_workspace_bytes(C.byref(d))
    if nbytes < 0:
        raise RuntimeError('essb_wgrad_workspace_bytes: unsupported segment layout')
    nbytes = max(int(nbytes), 4 * 1024 * Cout)
    ws = torch.empty((nbytes // 4 + 4,), device=dev, dtype=torch.float32)
    d.workspace, d.workspace_bytes = _p(ws), ws.numel() * 4
    call('essb_wgrad_fp32', C.byref(d), _stream())
    return dw, db


# ----------------------------------------------------------------------------- event pre-processing
def event_stats(data, T, Cw):
    """data [B, T*Cw, H, W] (contiguous per sample) -> stats [T, 3] double (sum, sumsq, nnz)."""
    B, _, H, W = data.shape
    if data.stride(1) != H * W or data.stride(2) != W or data.stride(3) != 1:
        raise RuntimeError('event tensor must be contiguous within each sample')
    stats = torch.empty((T, 3), device=data.device, dtype=torch.float64)
    call('essb_event_stats', _p(data), data.stride(0), B, T, Cw * H * W, _p(stats), _stream())
    return stats


def zero_pixels(data, xy):
    """Hot-pixel removal in place: data [B, C, H, W] (per-sample contiguous), xy int32 [n, 2] (x, y) on the device."""
    B, Cc, H, W = data.shape
    if data.stride(1) != H * W or data.stride(2) != W or data.stride(3) != 1:
        raise RuntimeError('event tensor must be contiguous within each sample')
    call('essb_zero_pixels', _p(data), data.stride(0), B, Cc, H, W, _p(xy), xy.shape[0], _stream())
    return data


def event_prepare(window, stats_row, normalize, Hp, Wp, pad_top, pad_left, ld_out, out=None, flip=False):
    """window: [B, C, H, W] view (per-sample contiguous) -> NHWC [B, Hp, Wp, ld_out] normalised+padded."""
    B, Cw, H, W = window.shape
    if window.stride(1) != H * W or window.stride(2) != W or window.stride(3) != 1:
        raise RuntimeError('event window must be contiguous within each sample')
    if out is None:
        out = torch.empty((B, Hp, Wp, ld_out), device=window.device, dtype=torch.float32)
    call('essb_event_prepare', _p(window), window.stride(0), _p(stats_row), int(normalize), _p(out), ld_out, B, Cw, H,
         W, Hp, Wp, pad_top, pad_left, int(flip), _stream())
    return out


HEAD_PAD_BEFORE, HEAD_PAD_AFTER_Y, HEAD_PAD_AFTER_X = 2, 2, 6


def head_planes_alloc(B, Hp, Wp, cpad, device):
    """Zero-bordered bf16 hi/lo buffers [B, Hp+4, Wp+8, cpad] for the tensor-core head convolution."""
    shape = (B, Hp + HEAD_PAD_BEFORE + HEAD_PAD_AFTER_Y, Wp + HEAD_PAD_BEFORE + HEAD_PAD_AFTER_X, cpad)
    return (torch.zeros(shape, device=device, dtype=torch.bfloat16), torch.zeros(shape, device=device, dtype=torch.bfloat16))


def event_prepare_planes(window, stats_row, normalize, Hp, Wp, pad_top, pad_left, planes, flip=False):
    """Like event_prepare but writes bf16 hi/lo planes into the interior of a head_planes_alloc buffer."""
    B, Cw, H, W = window.shape
    if window.stride(1) != H * W or window.stride(2) != W or window.stride(3) != 1:
        raise RuntimeError('event window must be contiguous within each sample')
    hi, lo = planes
    call('essb_event_prepare_planes', _p(window), window.stride(0), _p(stats_row), int(normalize), _p(hi), _p(lo),
         hi.shape[3], B, Cw, H, W, Hp, Wp, pad_top, pad_left, hi.shape[1], hi.shape[2], HEAD_PAD_BEFORE,
         HEAD_PAD_BEFORE, int(flip), _stream())
    return planes


def window_view(v: TcView, hi, lo, Hp, Wp, group=1):
    """Overlapping-stride TMA view of a head_planes_alloc buffer: view pixel (x, y) is the 8-pixel x cpad
    row window starting at padded pixel (group*x, y), i.e. image pixels (group*x-2 .. group*x+5, y-2) for
    tap row 0.  group > 1 makes one view pixel serve `group` adjacent output pixels."""
    N, Hb, Wb, cpad = hi.shape
    v.hi, v.lo = _p(hi), _p(lo)
    v.stride_x, v.stride_y, v.stride_n = group * cpad, Wb * cpad, Hb * Wb * cpad
    v.C, v.W, v.H = 8 * cpad, Wp // group, Hb


def nchw_to_nhwc(x, ld_out=None):
    """[N, C, H, W] (any strides) -> pixel-major [N, H, W, ld]; zero-copy for channels_last inputs."""
    N, Cc, H, W = x.shape
    x = x.float()
    v = x.permute(0, 2, 3, 1)
    if (ld_out is None or ld_out == Cc) and v.is_contiguous():
        return v
    x = x.contiguous()
    ld = ld_out or Cc
    out = (torch.zeros if ld != Cc else torch.empty)((N, H, W, ld), device=x.device, dtype=torch.float32)
    call('essb_nchw_to_nhwc', _p(x), _p(out), ld, N, Cc, H * W, _stream())
    return out


def as_nchw(t):
    """Pixel-major [N, H, W, C] buffer -> logical NCHW view (channels_last strides, zero copy)."""
    return t.permute(0, 3, 1, 2)


def bilinear_up2(x):
    N, H, W, Cc = x.shape
    out = torch.empty((N, 2 * H, 2 * W, Cc), device=x.device, dtype=torch.float32)
    call('essb_bilinear_up2', _p(x), _p(out), N, H, W, Cc, _stream())
    return out


# ----------------------------------------------------------------------------------------- task loss
def task_loss_sums(logits, target, K, ignore_index):
    """logits pixel-major [N, H, W, ld>=K]; target int64 [N, H, W] -> sums [2 + 3K] double."""
    npix = target.numel()
    sums = torch.empty((2 + 3 * K,), device=logits.device, dtype=torch.float64)
    call('essb_task_loss_fwd', _p(logits), logits.shape[-1], _p(target), npix, K, ignore_index, _p(sums), _stream())
    return sums


def task_loss_finish(sums, K, ignore_index, use_dice, use_ce):
    loss = torch.empty((1,), device=sums.device, dtype=torch.float32)
    call('essb_task_loss_finish', _p(sums), K, ignore_index, int(use_dice), int(use_ce), _p(loss), _stream())
    return loss


def task_loss_bwd(logits, target, K, ignore_index, sums, use_dice, use_ce, gscale):
    dl = torch.empty(tuple(logits.shape[:3]) + (K,), device=logits.device, dtype=torch.float32)
    call('essb_task_loss_bwd', _p(logits), logits.shape[-1], _p(target), target.numel(), K, ignore_index, _p(sums),
         int(use_dice), int(use_ce), _p(gscale), _p(dl), K, _stream())
    return dl


def confusion_labels(pred, target, K, ignore_index, conf=None):
    if conf is None:
        conf = torch.zeros((K, K), device=pred.device, dtype=torch.int64)
    call('essb_confusion_labels', _p(pred), _p(target), target.numel(), K, ignore_index, _p(conf), _stream())
    return conf


def confusion_logits(logits, target, K, ignore_index, conf=None):
    if conf is None:
        conf = torch.zeros((K, K), device=logits.device, dtype=torch.int64)
    call('essb_confusion', _p(logits), logits.shape[-1], _p(target), target.numel(), K, ignore_index, _p(conf),
         _stream())
    return conf


# --------------------------------------------------------------------------- tensor-core path helpers
PLANES_BF16, PLANES_HF8 = 0, 2     # operand plane formats (include/ess_b200.h: essb_split_planes)
PASSES = {'bf16': 1, 'f16f8': 2, 'bf16x3': 3}


def split_bf16(seg: Seg, N, H, W, hi=None, lo=None, c_off=0, ld_out=None, c_pad=0, fmt=PLANES_BF16):
    """fp32 (optionally normalised / ReLU'd / upsampled) -> tensor-core operand planes [N, H, W, ld_out]; channels
    [C, c_pad) of the output are zero-filled by the same kernel (K padding of a tensor-core operand).
    fmt PLANES_BF16: bf16 hi/lo; PLANES_HF8: fp16 hi + e4m3 pair plane (the f16f8 mode's operands; ld_out % 64 == 0).
    Both planes are carried as bfloat16-typed tensors of the same shape (2 bytes per element either way)."""
    s = Src()
    c = seg.fill(s)
    if hi is None:
        ld_out = ld_out or max(c, c_pad)
        hi = torch.empty((N, H, W, ld_out), device=seg.t.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
    call('essb_split_planes', C.byref(s), N, H, W, _p(hi), _p(lo), hi.shape[-1], c_off, c_pad, fmt, _stream())
    return hi, lo


def decode_planes(hi, lo, fmt=PLANES_BF16):
    """Operand planes -> the fp32 values they represent (diagnostics / tests; plain torch ops)."""
    if fmt == PLANES_BF16:
        return hi.float() + lo.float()
    ld = hi.shape[-1]
    x = hi.view(torch.float16).float() / 64.0
    b = lo.view(torch.uint8).reshape(tuple(lo.shape[:-1]) + (ld // 64, 128))
    a8l = b[..., 64:].contiguous().view(torch.float8_e4m3fn).float().reshape(tuple(lo.shape[:-1]) + (ld,))
    return x + a8l / float(1 << 14)


def pack_weight_tc(w, scale=None, transposed_layout=False, swap_io=False, flip=False, interleave=1, kin_pad=None,
                   nout_pad=None):
    """Reference-layout conv weight -> K-major bf16 hi/lo [NoutP][T*KinP] (essb_pack_weight_tc)."""
    w = w.detach().contiguous().float()
    if transposed_layout:
        cin, cout = w.shape[0], w.shape[1]
    else:
        cout, cin = w.shape[0], w.shape[1]
    T = w.shape[2] * w.shape[3]
    kin, nout = (cout, cin) if swap_io else (cin, cout)
    kinp = kin_pad or (kin + 63) // 64 * 64
    noutp = nout_pad or nout
    hi = torch.empty((noutp, T * kinp), device=w.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    call('essb_pack_weight_tc', _p(w), _p(scale), _p(hi), _p(lo), cout, cin, T, int(transposed_layout), int(swap_io),
         int(flip), int(interleave), kinp, noutp, _stream())
    return hi, lo, kinp


def pack_weight_tc_hf8(w, scale=None, transposed_layout=False, interleave=1, kin_pad=None, nout_pad=None):
    """Conv weight -> hf8 operand planes of the f16f8 mode (essb_pack_weight_tc_fmt, fmt 2).  Returns
    (hi, lo, KinP, acc_scale): the e4m3 scale 2^w8 is chosen so that max|W| (after the BN fold) lands in (64, 128];
    acc_scale = 2^-(w8+14) must be handed to the convolution that uses these weights.  One host sync (max) per pack --
    weights are packed once per parameter version."""
    w = w.detach().contiguous().float()
    if transposed_layout:
        cin, cout = w.shape[0], w.shape[1]
    else:
        cout, cin = w.shape[0], w.shape[1]
    ws = w
    if scale is not None:
        ws = w * (scale.view(1, -1, 1, 1) if transposed_layout else scale.view(-1, 1, 1, 1))
    m = float(ws.abs().max())
    w8 = int(math.floor(math.log2(128.0 / m))) if m > 0 and math.isfinite(m) else 0
    T = w.shape[2] * w.shape[3]
    kinp = kin_pad or (cin + 63) // 64 * 64
    noutp = nout_pad or cout
    hi = torch.empty((noutp, T * kinp), device=w.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    call('essb_pack_weight_tc_fmt', _p(w), _p(scale), _p(hi), _p(lo), cout, cin, T, int(transposed_layout), 0, 0,
         int(interleave), kinp, noutp, PLANES_HF8, w8, _stream())
    return hi, lo, kinp, 2.0 ** -(w8 + 14)


def _check_plane_strides(hi, lo):
    """Operand planes are dense inside an image ([H, W, ld] contiguous); only the batch stride may be larger."""
    N, H, W, ld = hi.shape
    for t in (hi, lo):
        if t is not None and (t.stride(3) != 1 or t.stride(2) != ld or t.stride(1) != W * ld or
                              (N > 1 and t.stride(0) < H * W * ld)):
            raise RuntimeError('operand planes must be pixel-major and dense inside each image (strides %s for shape %s)'
                               % (tuple(t.stride()), tuple(t.shape)))
    if lo is not None and (tuple(lo.shape) != tuple(hi.shape) or lo.stride(0) != hi.stride(0)):
        raise RuntimeError('hi / lo operand planes must share shape and strides')


def dense_view(v: TcView, hi, lo, c_off=0, C_=None):
    """TMA view of dense NHWC bf16 planes [N, H, W, ld]."""
    N, H, W, ld = hi.shape
    _check_plane_strides(hi, lo)
    v.hi, v.lo = _p(hi, c_off), _p(lo, c_off)
    v.stride_x, v.stride_y, v.stride_n = ld, W * ld, hi.stride(0)     # images may sit in a taller buffer (row-stacked levels)
    v.C, v.W, v.H = (C_ if C_ is not None else ld - c_off), W, H


def parity_view(v: TcView, hi, lo, py, px, fold_x=False):
    """View of the (py, px) parity plane of NHWC planes [N, H, W, ld] for stride-2 convolutions.
    fold_x: treat horizontal pixel pairs as one 2*ld-channel super-pixel (for ld = 32)."""
    N, H, W, ld = hi.shape
    _check_plane_strides(hi, lo)
    if fold_x:
        off = py * W * ld
        v.hi, v.lo = _p(hi, off), _p(lo, off)
        v.stride_x, v.stride_y, v.stride_n = 2 * ld, 2 * W * ld, hi.stride(0)
        v.C, v.W, v.H = 2 * ld, W // 2, H // 2
    else:
        off = (py * W + px) * ld
        v.hi, v.lo = _p(hi, off), _p(lo, off)
        v.stride_x, v.stride_y, v.stride_n = 2 * ld, 2 * W * ld, hi.stride(0)
        v.C, v.W, v.H = ld, W // 2, H // 2


def pick_bw_log2(OW, OH):
    """Spatial tile BW x (128/BW) (BW a power of two) minimising the padded pixel count."""
    best, best_waste = 4, None
    for b in (4, 3, 5, 6, 2):
        bw, bh = 1 << b, 128 >> b
        waste = ((OW + bw - 1) // bw * bw) * ((OH + bh - 1) // bh * bh) / float(OW * OH)
        if best_waste is None or waste < best_waste - 1e-9:
            best, best_waste = b, waste
    return best


def in_stats(y, count=None):
    """InstanceNorm statistics (mean, rstd) of a pixel-major tensor [N, H, W, C] (separate pass)."""
    N, H, W, Cc = y.shape
    blocks = _lib.lib().essb_in_bwd_blocks(H * W)
    partial = torch.empty((N, blocks, Cc, 2), device=y.device, dtype=torch.float32)
    call('essb_in_stats', _p(y), Cc, _p(partial), N, H * W, Cc, _stream())
    return in_finalize(partial, count or H * W)


def conv_tc_dense(planes, w_hi, w_lo, k_per_tap, taps, N, H, W, Cout, passes, bias=None, out=None, act=ACT_NONE,
                  tag=None, res_pre=None, res_post=None, out_planes=None, want_out=True, out_place=None, acc_scale=0.0,
                  planes_fmt=PLANES_BF16, phase_cout=0, flops=None):
    """Stride-1 gather-convolution on the tcgen05 kernel over a dense operand plane pair [N, H, W, Cin] (Cin a
    multiple of 64), or over a LIST of two such pairs = the channel concatenation of two tensors (each a
    multiple of 64 channels; alternatively concatenated inputs are laid out side by side by split_bf16(c_off=...)).  Returns the fp32 result [N, H, W, Cout] (None when want_out=False and
    only the bf16 `out_planes` are produced).  out_place = (OHf, OWf, osy, ooy, osx, oox) scatters the
    result into a larger image (transposed-conv phases).  phase_cout > 0: the four phases merged into ONE launch of
    Cout = 4 * phase_cout GEMM columns (merge_convT_phases; `out` / `out_planes` / residuals are phase_cout wide);
    `flops` = algorithmic FLOPs for the profile when they differ from the executed ones."""
    segs = list(planes) if isinstance(planes, list) else [planes]
    if not 1 <= len(segs) <= 2:
        raise ValueError('conv_tc_dense: one or two channel segments')
    hi, lo = segs[0]
    d = ConvTc()
    koff = 0
    for s, (shi, slo) in enumerate(segs):       # concatenated inputs (torch.cat on dim 1) = K segments, no copy
        dense_view(d.views[s], shi, slo)
        d.seg_C[s], d.seg_view0[s], d.seg_koff[s] = shi.shape[-1], s, koff
        koff += shi.shape[-1]
    d.n_views = d.nseg = len(segs)
    d.k_per_tap, d.n_w_taps, d.w_rows = k_per_tap, w_hi.shape[1] // k_per_tap, w_hi.shape[0]
    d.w_hi, d.w_lo, d.bias = _p(w_hi), _p(w_lo), _p(bias)
    if out_place is None:
        out_place = (H, W, 1, 0, 1, 0)
    if out is None and want_out:
        out = torch.empty((N, out_place[0], out_place[1], phase_cout or Cout), device=hi.device, dtype=torch.float32)
    if out is not None:
        d.out, d.ldo = _p(out), out.shape[-1]
    if out_planes is not None:
        d.out_hi, d.out_lo, d.ld_planes = _p(out_planes[0]), _p(out_planes[1]), out_planes[0].shape[-1]
    d.res_pre, d.res_post = _p(res_pre), _p(res_post)
    r = res_pre if res_pre is not None else res_post
    d.ld_res = r.shape[-1] if r is not None else 0
    d.N, d.OH, d.OW, d.Cout = N, H, W, Cout
    d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = out_place
    d.epilogue, d.act, d.passes, d.bw_log2 = EPI_LINEAR, act, passes, pick_bw_log2(W, H)
    d.acc_scale, d.planes_fmt, d.phase_cout = acc_scale, planes_fmt, phase_cout
    d.ntaps = len(taps)
    for t, (dy, dx, wi) in enumerate(taps):
        d.dy[t], d.dx[t], d.view[t], d.widx[t] = dy, dx, 0, wi
    conv_tc(d, tag=tag, device=hi.device, flops=flops)
    return out


def merge_convT_phases(wt):
    """ConvTranspose2d(k5, s2, p2, output_padding 1) weight [Cin, Cout, 5, 5] -> the weight [4*Cout, Cin, 3, 3] of ONE
    3x3 gather-convolution over the input whose output-channel block j holds sub-pixel phase (j >> 1, j & 1)
    (taps_convT_phase); the 11 of 36 (phase, offset) pairs a phase does not use are zero."""
    wt = wt.detach().float()
    cin, cout = wt.shape[0], wt.shape[1]
    wm = torch.zeros((4 * cout, cin, 3, 3), device=wt.device, dtype=torch.float32)
    for py in range(2):
        for px in range(2):
            ph = py * 2 + px
            for (dy, dx, widx) in taps_convT_phase(py, px):
                ky, kx = divmod(widx, 5)
                wm[ph * cout:(ph + 1) * cout, :, dy + 1, dx + 1] = wt[:, :, ky, kx].t()
    return wm


def wgrad_tc(a_planes, g_planes, Cin, Cout, taps, N, H, W, passes, tag='seg_wgrad', stride=1):
    """essb_wgrad_tc_run: dW [Cout, Cin, ntaps] from the bf16 planes of the conv input and of dY.  N, H, W are
    dY's dims; stride=2: the input planes are [N, 2H, 2W, Cin] and the taps' (dy, dx) are input-pixel offsets."""
    d = _lib.WgradTc()
    d.a_stride = stride
    d.a_hi, d.a_lo, d.g_hi, d.g_lo = _p(a_planes[0]), _p(a_planes[1]), _p(g_planes[0]), _p(g_planes[1])
    d.a_ld, d.Cin, d.g_ld, d.Cout = a_planes[0].shape[-1], Cin, g_planes[0].shape[-1], Cout
    d.N, d.H, d.W, d.passes = N, H, W, passes
    d.ntaps = len(taps)
    for i, tp in enumerate(taps):
        d.dy[i], d.dx[i] = tp[0], tp[1]
    dev = a_planes[0].device
    dw = torch.empty((Cout, Cin, len(taps)), device=dev, dtype=torch.float32)
    nbytes = _lib.lib().essb_wgrad_tc_workspace_bytes(C.byref(d))
    if nbytes < 0:
        raise RuntimeError('essb_wgrad_tc: unsupported shape')
    ws = torch.empty((int(nbytes) // 4 + 4,), device=dev, dtype=torch.float32)
    d.dw, d.workspace, d.workspace_bytes = _p(dw), _p(ws), ws.numel() * 4
    prof = _lib.PROFILE
    if prof is None or (_lib.PROFILE_TAGS is not None and tag not in _lib.PROFILE_TAGS):
        call('essb_wgrad_tc_run', C.byref(d), _stream())
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call('essb_wgrad_tc_run', C.byref(d), _stream())
        e1.record()
        prof.append((tag, 2.0 * N * H * W * Cout * Cin * len(taps), e0, e1))
    return dw


def dgrad_phase_taps(k, pad, stride):
    """Input-gradient of a k x k / stride s / padding `pad` convolution as gather-convs over dY, one per
    output phase:  dx[s*i'+p] = sum_{k = p+pad (mod s)} dy[i' + (p+pad-k)/s] * w[k].
    Returns {(py, px): [(dy, dx, widx), ...]} (an empty list = that phase is identically zero)."""
    out = {}
    for py in range(stride):
        for px in range(stride):
            taps = []
            for ky in range(k):
                if (py + pad - ky) % stride:
                    continue
                for kx in range(k):
                    if (px + pad - kx) % stride:
                        continue
                    taps.append(((py + pad - ky) // stride, (px + pad - kx) // stride, ky * k + kx))
            out[(py, px)] = taps
    return out


def bn_train_finalize(mean, rstd, gamma, beta, rows, bn, momentum, eps):
    """(a, b) of the BN apply pass; updates bn.running_mean / running_var / num_batches_tracked in place when `bn`
    (an nn.BatchNorm2d with fp32 running statistics on this device) is given.  One launch (essb_bn_train_finalize)."""
    Cc = mean.numel()
    a = torch.empty((Cc,), device=mean.device, dtype=torch.float32)
    b = torch.empty_like(a)
    rm = rv = nb = None
    if bn is not None:
        rm, rv, nb = bn.running_mean, bn.running_var, bn.num_batches_tracked
        if rm.dtype != torch.float32 or rv.dtype != torch.float32 or rm.device != mean.device or not rm.is_contiguous() \
                or not rv.is_contiguous():
            raise RuntimeError('bn_train_finalize: running statistics must be contiguous fp32 on the activation\'s device')
        if nb is not None and (nb.dtype != torch.int64 or nb.device != mean.device):
            nb += 1           # exotic placement: keep torch semantics
            nb = None
    call('essb_bn_train_finalize', _p(mean), _p(rstd), _p(gamma), _p(beta), _p(a), _p(b), _p(rm), _p(rv), _p(nb), Cc,
         int(rows), float(momentum), float(eps), _stream())
    if bn is not None:        # raw-pointer writes: keep autograd's / state_dict consumers' version counters honest
        torch._C._increment_version([t for t in (rm, rv, nb) if t is not None])
    return a, b


def affine_act(x, a, b, res=None, relu=False, out=None):
    """out = relu?(x * a[c] + b[c] + res) over a pixel-major tensor [..., C]."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    if out is None:
        out = torch.empty_like(x)
    call('essb_affine_act', _p(x), Cc, _p(a), _p(b), _p(res), Cc if res is not None else 0, int(relu), _p(out), Cc,
         rows, Cc, _stream())
    return out


def bn_backward(dout, mask_src, x, mean, rstd, gamma):
    """Train-mode BatchNorm backward.  Returns (dx, dgamma, dbeta, g) with g = dout * (mask_src > 0)."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    dev = x.device
    blocks = _lib.lib().essb_bn_bwd_blocks(rows)
    g = torch.empty_like(x)
    partial = torch.empty((1, blocks, Cc, 2), device=dev, dtype=torch.float32)
    call('essb_bn_bwd_pass1', _p(dout), Cc, _p(mask_src), Cc if mask_src is not None else 0, _p(x), Cc, _p(mean),
         _p(rstd), _p(g), _p(partial), rows, Cc, _stream())
    totals = torch.empty((1, Cc, 2), device=dev, dtype=torch.float32)
    call('essb_partial_reduce', _p(partial), 1, blocks, Cc, _p(totals), _stream())
    dx = torch.empty_like(x)
    call('essb_bn_bwd_pass2', _p(g), _p(x), Cc, _p(mean), _p(rstd), _p(gamma), _p(totals), _p(dx), rows, Cc, _stream())
    return dx, totals[0, :, 1].contiguous(), totals[0, :, 0].contiguous(), g


def l1_sum(a, b):
    sums = torch.empty((1,), device=a.device, dtype=torch.float64)
    call('essb_l1_fwd', _p(a), _p(b), a.numel(), _p(sums), _stream())
    return sums


def l1_bwd(a, b, gscale):
    da = torch.empty_like(a)
    call('essb_l1_bwd', _p(a), _p(b), a.numel(), _p(gscale), _p(da), _stream())
    return da


def jsdiv(predict, target, K, want_sum=True, gscale=None, want_grad=False):
    """symJSDivLoss kernels on pixel-major logits [N, H, W, ld]; returns (sums | None, dpredict | None)."""
    rows = predict.numel() // predict.shape[-1]
    sums = torch.empty((1,), device=predict.device, dtype=torch.float64) if want_sum else None
    dp = torch.empty(tuple(predict.shape[:-1]) + (K,), device=predict.device, dtype=torch.float32) if want_grad else None
    call('essb_jsdiv', _p(predict), predict.shape[-1], _p(target), target.shape[-1], rows, K, _p(sums), _p(gscale),
         _p(dp), K, _stream())
    return sums, dp


def pw_conv_fwd(seg: Seg, w, bias, N, H, W, Cout, act=ACT_NONE):
    """essb_pw_conv_fwd: 1x1 conv (Cin 32/64 -> Cout <= 16) with the source's IN/ReLU applied on the fly.
    w [Cout, Cin] fp32 contiguous (reference layout)."""
    s = Src()
    seg.fill(s)
    out = torch.empty((N, H, W, Cout), device=seg.t.device, dtype=torch.float32)
    call('essb_pw_conv_fwd', C.byref(s), _p(w), _p(bias), _p(out), Cout, N, H, W, Cout, act, _stream())
    return out


def pw_conv_dgrad(dy, w, Cin):
    """essb_pw_conv_dgrad: dy [N, H, W, Cout], w [Cout, Cin] -> gradient w.r.t. the conv input [N, H, W, Cin]."""
    N, H, W, Cout = dy.shape
    dx = torch.empty((N, H, W, Cin), device=dy.device, dtype=torch.float32)
    call('essb_pw_conv_dgrad', _p(dy), Cout, _p(w), _p(dx), Cin, N * H * W, Cin, Cout, _stream())
    return dx


def pw_conv_wgrad(seg: Seg, dy, want_w=True, want_b=True):
    """essb_pw_conv_wgrad -> (dW [Cout, Cin], dbias [Cout]) (None where not wanted)."""
    N, H, W, Cout = dy.shape
    s = Src()
    cin = seg.fill(s)
    dev = dy.device
    dw = torch.empty((Cout, cin), device=dev, dtype=torch.float32) if want_w else None
    db = torch.empty((Cout,), device=dev, dtype=torch.float32) if want_b else None
    nbytes = _lib.lib().essb_pw_conv_wgrad_workspace_bytes(N, H, W, cin)
    if nbytes < 0:
        raise RuntimeError('essb_pw_conv_wgrad: unsupported shape')
    ws = torch.empty((int(nbytes) // 4,), device=dev, dtype=torch.float32)
    call('essb_pw_conv_wgrad', C.byref(s), _p(dy), Cout, N, H, W, Cout, _p(dw), _p(db), _p(ws), ws.numel() * 4, _stream())
    return dw, db


def stem_conv_supported(cin, cout, k, stride):
    return cin == 1 and bool(_lib.lib().essb_stem_conv_supported(cout, k, stride))


def stem_conv_fwd(x, w, stride, pad):
    """essb_stem_conv_fwd: x [N, H, W, 1] fp32, w [Cout, 1, k, k] -> [N, OH, OW, Cout]."""
    N, H, W, _ = x.shape
    Cout, _, k, _ = w.shape
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty((N, OH, OW, Cout), device=x.device, dtype=torch.float32)
    call('essb_stem_conv_fwd', _p(x), _p(w), _p(out), N, H, W, Cout, k, stride, pad, _stream())
    return out


def stem_conv_wgrad(x, dy, Cout, k, stride, pad):
    """essb_stem_conv_wgrad -> dW [Cout, 1, k, k]."""
    N, H, W, _ = x.shape
    dw = torch.empty((Cout, 1, k, k), device=x.device, dtype=torch.float32)
    nbytes = _lib.lib().essb_stem_conv_wgrad_workspace_bytes(Cout, k)
    if nbytes < 0:
        raise RuntimeError('essb_stem_conv_wgrad: unsupported shape')
    ws = torch.empty((int(nbytes) // 4,), device=x.device, dtype=torch.float32)
    call('essb_stem_conv_wgrad', _p(x), _p(dy), _p(dw), N, H, W, Cout, k, stride, pad, _p(ws), ws.numel() * 4, _stream())
    return dw


def colsum(x, Cc=None):
    """Column sums over all pixels of a pixel-major tensor [N, H, W, ld] -> [C] (bias gradients)."""
    rows = x.shape[0] * x.shape[1] * x.shape[2]
    Cc = Cc or x.shape[-1]
    out = torch.empty((Cc,), device=x.device, dtype=torch.float32)
    ws = torch.empty((1024 * Cc,), device=x.device, dtype=torch.float32)
    call('essb_colsum', _p(x), x.shape[-1], rows, Cc, _p(out), _p(ws), ws.numel() * 4, _stream())
    return out


def conv_tc_s2(planes, w_hi, w_lo, k_per_tap, k, pad, N, H, W, Cout, passes, bias=None, act=ACT_NONE, tag=None):
    """k x k stride-2 convolution on the tcgen05 kernel: the input planes [N, H, W, Cin] (Cin a multiple of
    64, H and W even) are addressed through four parity views; tap offset k-pad = 2a + p selects view p and
    plane offset a.  Returns the fp32 result [N, H/2, W/2, Cout]."""
    hi, lo = planes
    d = ConvTc()
    for py in range(2):
        for px in range(2):
            parity_view(d.views[py * 2 + px], hi, lo, py, px)
    d.n_views, d.nseg = 4, 1
    d.seg_C[0], d.seg_view0[0], d.seg_koff[0] = hi.shape[-1], 0, 0
    d.k_per_tap, d.n_w_taps, d.w_rows = k_per_tap, k * k, w_hi.shape[0]
    d.w_hi, d.w_lo, d.bias = _p(w_hi), _p(w_lo), _p(bias)
    OH, OW = H // 2, W // 2
    out = torch.empty((N, OH, OW, Cout), device=hi.device, dtype=torch.float32)
    d.out, d.ldo = _p(out), Cout
    d.N, d.OH, d.OW, d.Cout = N, OH, OW, Cout
    d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = OH, OW, 1, 0, 1, 0
    d.epilogue, d.act, d.passes, d.bw_log2 = EPI_LINEAR, act, passes, pick_bw_log2(OW, OH)
    t = 0
    for ky in range(k):
        for kx in range(k):
            oy, ox = ky - pad, kx - pad
            py, px = oy % 2, ox % 2
            d.dy[t], d.dx[t], d.view[t], d.widx[t] = (oy - py) // 2, (ox - px) // 2, py * 2 + px, ky * k + kx
            t += 1
    d.ntaps = t
    conv_tc(d, tag=tag, device=hi.device)
    return out


_TC_POOL = {}
TC_SCHED_SLOTS = 256
TC_SPLITK_WS_BYTES = 80 << 20      # covers 7 K-slices of the 84 tail tiles of the 1/8-scale ConvLSTM at B=8 (75 MB)
# Split-K of the tail wave is OFF by default: measured on B200 (profiles/r01_bench_g_splitk.json) the fused
# ConvLSTM kernel is power-bound, not tail-bound (SM clock drops to ~1.54 GHz under the 1 kW cap while the idle SMs
# of the last wave give their power budget to the busy ones), so the extra partial-accumulator traffic costs
# more (+7.7 % kernel time) than the recovered wave quantisation.  ESSB_TC_SPLITK=1 turns it on.
SPLITK = os.environ.get('ESSB_TC_SPLITK', '0') == '1'


def _tc_workspace(device):
    """Per (device, stream) scratch of the tcgen05 launches: rotating scheduler slots (2 zeroed int32 each,
    left zero by every launch), split-K arrival counters and the split-K partial-accumulator buffer.
    Launches on one stream are ordered, so they can share the buffers."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ent = _TC_POOL.get(key)
    if ent is None:
        ent = {'sched': torch.zeros((TC_SCHED_SLOTS, 2), device=device, dtype=torch.int32), 'next': 0,
               'cnt': torch.zeros((148 * 8,), device=device, dtype=torch.int32), 'ws': None}
        _TC_POOL[key] = ent
    if SPLITK and ent['ws'] is None:     # the 80 MB partial-accumulator buffer exists only when split-K is switched on
        ent['ws'] = torch.empty((TC_SPLITK_WS_BYTES // 4,), device=device, dtype=torch.float32)
    return ent


def conv_tc(d: ConvTc, tag=None, device=None, flops=None):
    """essb_conv_tc_run; when _lib.PROFILE is a list, brackets the launch with CUDA events on the
    launching stream and records (tag, algorithmic FLOPs, start, end).  `device`: the device the operands live
    on (scheduler slots / split-K scratch are allocated there; default = the current device)."""
    if device is None or device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    ent = _tc_workspace(device)
    slot = ent['next']
    ent['next'] = (slot + 1) % TC_SCHED_SLOTS
    d.sched = _p(ent['sched'], 2 * slot)
    if SPLITK:
        d.splitk_ws, d.splitk_cnt, d.splitk_ws_bytes = _p(ent['ws']), _p(ent['cnt']), TC_SPLITK_WS_BYTES
    prof = _lib.PROFILE
    if prof is None or (_lib.PROFILE_TAGS is not None and tag not in _lib.PROFILE_TAGS):
        call('essb_conv_tc_run', C.byref(d), _stream())
        return
    k = sum(d.seg_C[s] for s in range(d.nseg)) * d.ntaps
    if flops is None:
        flops = 2.0 * d.N * d.OH * d.OW * d.Cout * k
    # under CUDA-graph capture the events become event-record NODES (external=True): after every replay they hold the
    # timestamps of that replay, so a profile taken at capture time keeps measuring (entries carry a 5th field True)
    cap = torch.cuda.is_current_stream_capturing()
    e0, e1 = (torch.cuda.Event(enable_timing=True, external=cap), torch.cuda.Event(enable_timing=True, external=cap))
    e0.record()
    call('essb_conv_tc_run', C.byref(d), _stream())
    e1.record()
    prof.append((tag or 'conv_tc', flops, e0, e1, True) if cap else (tag or 'conv_tc', flops, e0, e1))
