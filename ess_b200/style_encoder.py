"""Drop-in `StyleEncoderE2VID` (reference: models/style_networks.py:110-145) -- the UDA image encoder:
ResNet-18 stem (conv7x7 s2 + BN + ReLU, no max-pool) + layer1..3 with BatchNorm in TRAIN mode
(training/ess_trainer.py:159-162), forward and backward on libess_b200.so kernels.

Same constructor (`input_dim`, `skip_connect`), `forward(x) -> {1: x, 2, 4, 8}` and `state_dict` keys
(`encoder_scale_1.0.weight`, `encoder_scale_1.1.*`, `encoder_scale_1.3.{0,1}.*`, `encoder_scale_{2,3}.{0,1}.*`)
as the reference.  The torchvision modules are parameter/buffer holders only.  Like the reference
(`resnet18(pretrained=True)`, style_networks.py:117-121) the holders start from the ImageNet weights: taken from
the torch hub cache, else downloaded; when neither is possible (no network) a warning says so LOUDLY and the
encoder keeps torchvision's random init (ESS_B200_PRETRAINED=0 opts out silently, e.g. for checkpoint loading).

Execution: two small autograd Functions on pixel-major fp32 tensors -- a bias-free gather convolution
(forward; input gradient as per-phase gather-convs over dY; weight gradient as split-K implicit GEMM)
and train-mode BatchNorm fused with the BasicBlock residual add and ReLU (batch statistics in one
memory pass, apply in one pass; backward in two passes).  In the tensor-core modes every 3x3 / 1x1 convolution
runs on the tcgen05 kernels (stride 2 through parity views; input gradients of strided convs as per-phase
gathers with strided stores); the 1-channel 7x7 stem has its own HBM-bound kernels (stem_conv.cu); the stride-2
weight gradients read the input through parity planes on the tcgen05 wgrad kernel.
"""
import os
import warnings

import torch
import torch.nn as nn
import torchvision.models as tvm

from . import ops
from ._lib import ACT_NONE
from .ops import Seg

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _tc_ok(mode, k, cin, cout):
    return mode != 'fp32' and k in (1, 3) and cin % 64 == 0 and cout in (64, 128, 256)


class _ConvFn(torch.autograd.Function):
    """y = conv2d(x, w, stride, padding) without bias; x, y pixel-major [N, H, W, C].
    mode 'fp32': exact CUDA-core kernels; 'bf16x3' / 'bf16': tcgen05 kernels wherever the shape allows
    (3x3 / 1x1 with Cin % 64 == 0: forward, input gradient; weight gradient for stride 1)."""

    @staticmethod
    def forward(ctx, x, w, stride, pad, mode):
        N, H, W, Cin = x.shape
        Cout, _, k, _ = w.shape
        OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        passes = 1 if mode == 'bf16' else 3      # 'f16f8' (an E2VID mode) = bf16x3 here
        if _tc_ok(mode, k, Cin, Cout) and (stride == 1 or (stride == 2 and H % 2 == 0 and W % 2 == 0)):
            planes = ops.split_bf16(Seg(x), N, H, W)
            w_hi, w_lo, kinp = ops.pack_weight_tc(w)
            if stride == 1:
                y = ops.conv_tc_dense(planes, w_hi, w_lo, kinp, ops.taps_conv(k, pad), N, H, W, Cout, passes, tag='uda_fwd')
            else:
                y = ops.conv_tc_s2(planes, w_hi, w_lo, kinp, k, pad, N, H, W, Cout, passes, tag='uda_fwd')
        elif ops.stem_conv_supported(Cin, Cout, k, stride) and x.is_contiguous():
            y = ops.stem_conv_fwd(x, w.detach().float().contiguous(), stride, pad)      # 1-channel 7x7 stem (HBM-bound)
        else:
            wp = ops.pack_weight(w)
            y, _, _, _ = ops.conv([Seg(x)], wp, None, N, H, W, OH, OW, Cout, ops.taps_conv(k, pad), stride=stride,
                                  act=ACT_NONE)
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad, mode)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, pad, mode = ctx.cfg
        N, H, W, Cin = x.shape
        Cout, _, k, _ = w.shape
        gy = gy.contiguous()
        OH, OW = gy.shape[1], gy.shape[2]
        passes = 1 if mode == 'bf16' else 3      # 'f16f8' (an E2VID mode) = bf16x3 here
        tc = _tc_ok(mode, k, Cin, Cout)
        gplanes = ops.split_bf16(Seg(gy), N, OH, OW) if tc else None          # Cout is a multiple of 64
        gx = gw = None
        if ctx.needs_input_grad[1]:
            if tc and (stride == 1 or (stride == 2 and H == 2 * OH and W == 2 * OW)):
                planes = ops.split_bf16(Seg(x), N, H, W)
                gw = ops.wgrad_tc(planes, gplanes, Cin, Cout, ops.taps_conv(k, pad), N, OH, OW, passes,
                                  tag='uda_wgrad', stride=stride).view(w.shape)
            elif ops.stem_conv_supported(Cin, Cout, k, stride) and x.is_contiguous():
                gw = ops.stem_conv_wgrad(x, gy, Cout, k, stride, pad).view(w.shape)
            else:
                dw, _ = ops.wgrad([Seg(x)], gy, N, H, W, OH, OW, Cout, ops.taps_conv(k, pad), stride=stride,
                                  want_bias=False)
                gw = dw.view(w.shape)
        if ctx.needs_input_grad[0]:
            if H != OH * stride or W != OW * stride:
                raise RuntimeError('conv input gradient needs H, W divisible by the stride')
            phases = ops.dgrad_phase_taps(k, pad, stride)
            alloc = torch.zeros if any(not t for t in phases.values()) else torch.empty
            gx = alloc((N, H, W, Cin), device=x.device, dtype=torch.float32)
            if tc:
                w_hi, w_lo, kinp = ops.pack_weight_tc(w.detach(), swap_io=True)
                for (py, px), taps in phases.items():
                    if taps:
                        ops.conv_tc_dense(gplanes, w_hi, w_lo, kinp, taps, N, OH, OW, Cin, passes, out=gx,
                                          out_place=(H, W, stride, py, stride, px), tag='uda_dgrad')
            else:
                wp = ops.pack_weight(w.detach(), swap_io=True)
                for (py, px), taps in phases.items():
                    if taps:
                        ops.conv([Seg(gy)], wp, None, N, OH, OW, OH, OW, Cin, taps, out=gx,
                                 out_place=(H, W, stride, py, stride, px))
        return gx, gw, None, None, None


class _BNFn(torch.autograd.Function):
    """out = relu?(BatchNorm_train(x) + res): batch statistics over all N*H*W rows per channel."""

    @staticmethod
    def forward(ctx, x, gamma, beta, res, relu, bn):
        N, H, W, Cc = x.shape
        rows = N * H * W
        mean, rstd = ops.in_stats(x.view(1, 1, rows, Cc))          # biased variance, eps inside the sqrt
        mean, rstd = mean.view(-1).contiguous(), rstd.view(-1).contiguous()
        g = gamma.detach().float().contiguous()
        # folded scale/shift + nn.BatchNorm2d running-statistics update (unbiased variance, momentum form) +
        # num_batches_tracked: ONE launch (was ~12 ATen launches per BatchNorm layer)
        a, b = ops.bn_train_finalize(mean, rstd, g, beta.detach().float().contiguous(), rows,
                                     bn if (bn is not None and bn.track_running_stats) else None, BN_MOMENTUM, BN_EPS)
        out = ops.affine_act(x, a, b, res=res, relu=relu)
        ctx.save_for_backward(x, mean, rstd, g, out if relu else None)
        ctx.has_res = res is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        x, mean, rstd, gamma, out = ctx.saved_tensors
        dx, dgamma, dbeta, g = ops.bn_backward(gout.contiguous(), out, x, mean, rstd, gamma)
        return dx, dgamma, dbeta, (g if ctx.has_res else None), None, None


def _bn_apply(x, bn, res, relu, training):
    if training:
        return _BNFn.apply(x, bn.weight, bn.bias, res, relu, bn)
    if torch.is_grad_enabled() and (x.requires_grad or bn.weight.requires_grad):
        raise NotImplementedError('StyleEncoderE2VID in eval mode is forward-only (use torch.no_grad())')
    a = (bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + BN_EPS)).contiguous()
    b = (bn.bias.detach().float() - bn.running_mean.float() * a).contiguous()
    return ops.affine_act(x, a, b, res=res, relu=relu)


_warned_pretrained = [False]


def _resnet18_imagenet():
    """torchvision resnet18 with the ImageNet weights the reference starts from (style_networks.py:117-121).
    Returns (module, source) with source in {'cache', 'download', 'disabled', 'random-init'}."""
    if os.environ.get('ESS_B200_PRETRAINED', '1') == '0':
        return tvm.resnet18(weights=None), 'disabled'
    w = tvm.ResNet18_Weights.IMAGENET1K_V1
    ckpt = os.path.join(torch.hub.get_dir(), 'checkpoints', os.path.basename(w.url))
    try:
        if os.path.exists(ckpt):
            r = tvm.resnet18(weights=None)
            r.load_state_dict(torch.load(ckpt, map_location='cpu', weights_only=True))
            return r, 'cache'
        import socket
        old = socket.getdefaulttimeout()
        socket.setdefaulttimeout(10)
        try:
            return tvm.resnet18(weights=w, progress=False), 'download'
        finally:
            socket.setdefaulttimeout(old)
    except Exception as ex:
        if not _warned_pretrained[0]:
            _warned_pretrained[0] = True
            warnings.warn('ess_b200.StyleEncoderE2VID: ImageNet weights for resnet18 are NOT available (%s: %s); the '
                          'image encoder starts from RANDOM init, unlike the reference (resnet18(pretrained=True), '
                          'models/style_networks.py:117-121).  Put %s into %s, load a checkpoint, or set '
                          'ESS_B200_PRETRAINED=0 to acknowledge.' % (type(ex).__name__, str(ex)[:80],
                                                                      os.path.basename(w.url), os.path.dirname(ckpt)),
                          RuntimeWarning, stacklevel=3)
        return tvm.resnet18(weights=None), 'random-init'


class StyleEncoderE2VID(nn.Module):
    def __init__(self, input_dim, skip_connect=False):
        super().__init__()
        self.skip_connect = skip_connect
        r, self.pretrained_source = _resnet18_imagenet()
        conv_list = [nn.Conv2d(input_dim, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3), bias=False)]
        conv_list += list(r.children())[1:3]          # bn1, relu
        conv_list += list(r.children())[4:5]          # layer1 (max-pool skipped, as in the reference)
        self.encoder_scale_1 = nn.Sequential(*conv_list)
        self.encoder_scale_2 = list(r.children())[5]  # layer2
        self.encoder_scale_3 = list(r.children())[6]  # layer3
        from .e2vid import default_mode
        self.mode = default_mode()    # 'fp32' | 'bf16x3' | 'bf16' (same meaning as in E2VIDRecurrent / SemSegE2VID)

    def update_skip_dict(self, skips, x, sz_in):
        rem, scale = sz_in % x.shape[3], sz_in // x.shape[3]
        assert rem == 0
        skips[scale] = x

    def _block(self, x, blk):
        """torchvision BasicBlock: conv-bn-relu-conv-bn (+ downsample) + add + relu."""
        tr = self.training
        s = blk.conv1.stride[0]
        y = _ConvFn.apply(x, blk.conv1.weight, s, 1, self.mode)
        y = _bn_apply(y, blk.bn1, None, True, tr)
        y = _ConvFn.apply(y, blk.conv2.weight, 1, 1, self.mode)
        identity = x
        if blk.downsample is not None:
            identity = _ConvFn.apply(x, blk.downsample[0].weight, blk.downsample[0].stride[0], 0, self.mode)
            identity = _bn_apply(identity, blk.downsample[1], None, False, tr)
        return _bn_apply(y, blk.bn2, identity, True, tr)

    def forward(self, x):
        with ops.on_device_of(x):
            return self._forward(x)

    def _forward(self, x):
        out = {1: x}
        sz_in = x.shape[3]
        if x.shape[2] % 8 or x.shape[3] % 8:
            raise RuntimeError('StyleEncoderE2VID: H, W must be multiples of 8')
        t = x.float().permute(0, 2, 3, 1).contiguous()           # C = input_dim pixel-major
        e1 = self.encoder_scale_1
        y = _ConvFn.apply(t, e1[0].weight, 2, 3, self.mode)
        y = _bn_apply(y, e1[1], None, True, self.training)
        for blk in e1[3]:
            y = self._block(y, blk)
        if self.skip_connect:
            self.update_skip_dict(out, y.permute(0, 3, 1, 2), sz_in)
        for blk in self.encoder_scale_2:
            y = self._block(y, blk)
        if self.skip_connect:
            self.update_skip_dict(out, y.permute(0, 3, 1, 2), sz_in)
        for blk in self.encoder_scale_3:
            y = self._block(y, blk)
        self.update_skip_dict(out, y.permute(0, 3, 1, 2), sz_in)
        return out
