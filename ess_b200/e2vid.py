"""Drop-in `E2VIDRecurrent` (reference: e2vid/model/model.py:69-100 -> e2vid/model/unet.py:117-181).

Same constructor (`config` dict), attributes (`num_bins`, `num_encoders`, ...), `forward(event_tensor,
prev_states) -> (img, states, latent)` signature and `state_dict` keys as the reference, so
`e2vid.utils.loading_utils.load_model` (`eval(arch)(cfg)` + strict `load_state_dict`) and
`ImageReconstructor` work unchanged.  The nn.Conv2d / nn.BatchNorm2d children are parameter holders
only -- their forward is never called; all arithmetic runs in libess_b200.so:

  mode "fp32"   : every conv on the exact-fp32 CUDA-core implicit-GEMM kernel (conv_fp32.cu)
  mode "bf16x3" : encoder stride-2 convs and the fused ConvLSTM cells on the tcgen05/TMA kernel
                  (conv_tc.cu) with the 3-product bf16 split (fp32-level accuracy)
  mode "f16f8"  : same kernels, operands in the "hf8" format: fp16 main product + two e4m3 cross terms issued as
                  ONE K=128 fp8 product per 64-channel chunk -- 2 tensor-pass equivalents instead of 3 at
                  fp32-level accuracy (tools/precision_emul.py; the head conv stays bf16x3: its overlapping-window
                  operand view is incompatible with the paired fp8 layout)
  mode "bf16"   : same kernels, single bf16 pass (fast, ~1e-2 relative; reported separately)

Inference only (the reference freezes this network and runs it under no_grad,
training/ess_supervised_trainer.py:44-47, e2vid/image_reconstructor.py:83); BatchNorm is applied in
eval mode (running statistics folded into the conv weights).
"""
import os

import torch
import torch.nn as nn

from . import ops
from ._lib import (ACT_NONE, ACT_RELU, ACT_SIGMOID, EPI_GRU_OUT, EPI_GRU_UR, EPI_LINEAR, EPI_LSTM, ConvTc)
from .ops import Seg

BN_EPS = 1e-5
MODES = ('fp32', 'bf16x3', 'bf16', 'f16f8')


_WARNED = set()


def _warn_once(msg):
    """A configuration that leaves the tensor-core path is correct but ~10x slower: say so once per message."""
    if msg not in _WARNED:
        _WARNED.add(msg)
        import warnings
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


def default_mode():
    m = os.environ.get('ESS_B200_MODE', 'f16f8')
    if m not in MODES:
        raise ValueError('ESS_B200_MODE must be one of %s' % (MODES,))
    return m


# ---------------------------------------------------------------------------- parameter holders
class _ConvLayer(nn.Module):          # e2vid/model/submodules.py:7-31
    def __init__(self, cin, cout, k, stride, pad, norm):
        super().__init__()
        self.conv2d = nn.Conv2d(cin, cout, k, stride, pad, bias=(norm != 'BN'))
        if norm == 'BN':
            self.norm_layer = nn.BatchNorm2d(cout)
        elif norm == 'IN':                # submodules.py:21-22: running statistics, no affine parameters
            self.norm_layer = nn.InstanceNorm2d(cout, track_running_stats=True)


class _TransposedConvLayer(nn.Module):  # submodules.py:34-62
    def __init__(self, cin, cout, k, pad, norm):
        super().__init__()
        self.transposed_conv2d = nn.ConvTranspose2d(cin, cout, k, stride=2, padding=pad, output_padding=1,
                                                    bias=(norm != 'BN'))
        if norm == 'BN':
            self.norm_layer = nn.BatchNorm2d(cout)
        elif norm == 'IN':                # submodules.py:50-51
            self.norm_layer = nn.InstanceNorm2d(cout, track_running_stats=True)


class _ConvLSTM(nn.Module):           # submodules.py:175-188
    def __init__(self, c, hidden, k):
        super().__init__()
        self.Gates = nn.Conv2d(c + hidden, 4 * hidden, k, padding=k // 2)


class _ConvGRU(nn.Module):            # submodules.py:233-253
    def __init__(self, c, hidden, k):
        super().__init__()
        pad = k // 2
        self.reset_gate = nn.Conv2d(c + hidden, hidden, k, padding=pad)
        self.update_gate = nn.Conv2d(c + hidden, hidden, k, padding=pad)
        self.out_gate = nn.Conv2d(c + hidden, hidden, k, padding=pad)
        for m in (self.reset_gate, self.update_gate, self.out_gate):
            nn.init.orthogonal_(m.weight)
            nn.init.constant_(m.bias, 0.)


class _RecurrentConvLayer(nn.Module):  # submodules.py:96-115
    def __init__(self, cin, cout, block_type, norm):
        super().__init__()
        self.conv = _ConvLayer(cin, cout, 5, 2, 2, norm)
        self.recurrent_block = (_ConvLSTM if block_type == 'convlstm' else _ConvGRU)(cout, cout, 3)


class _ResidualBlock(nn.Module):      # submodules.py:140-155
    def __init__(self, c, norm):
        super().__init__()
        bias = norm != 'BN'
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1, bias=bias)
        if norm == 'BN':
            self.bn1 = nn.BatchNorm2d(c)
            self.bn2 = nn.BatchNorm2d(c)
        elif norm == 'IN':                # submodules.py:149-151: plain InstanceNorm2d (per-sample statistics, nothing stored)
            self.bn1 = nn.InstanceNorm2d(c)
            self.bn2 = nn.InstanceNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1, bias=bias)


class _UNetRecurrent(nn.Module):      # unet.py:117-143
    def __init__(self, cfg):
        super().__init__()
        base, ne, norm = cfg['base_num_channels'], cfg['num_encoders'], cfg['norm']
        concat = cfg['skip_type'] != 'sum'
        self.head = _ConvLayer(cfg['num_bins'], base, 5, 1, 2, None)
        self.encoders = nn.ModuleList(
            [_RecurrentConvLayer(base * 2 ** i, base * 2 ** (i + 1), cfg['recurrent_block_type'], norm)
             for i in range(ne)])
        cmax = base * 2 ** ne
        self.resblocks = nn.ModuleList([_ResidualBlock(cmax, norm) for _ in range(cfg['num_residual_blocks'])])
        self.decoders = nn.ModuleList()
        for c in reversed([base * 2 ** (i + 1) for i in range(ne)]):
            cin = 2 * c if concat else c
            if cfg['use_upsample_conv']:
                self.decoders.append(_ConvLayer(cin, c // 2, 5, 1, 2, norm))
            else:
                self.decoders.append(_TransposedConvLayer(cin, c // 2, 5, 2, norm))
        self.pred = _ConvLayer(2 * base if concat else base, 1, 1, 1, 0, norm)


def _bn_fold(conv_bias, bn, cout, device):
    """Eval-mode BatchNorm -> (scale, bias) per output channel (submodules.py:19-20,26-27)."""
    if bn is None or getattr(bn, 'running_var', None) is None:      # no norm, or a plain InstanceNorm2d (handled by its caller)
        return None, (conv_bias.detach().float().contiguous() if conv_bias is not None else None)
    if isinstance(bn, nn.InstanceNorm2d):      # norm='IN' conv layers: eval-mode InstanceNorm2d(track_running_stats=True)
        scale = 1.0 / torch.sqrt(bn.running_var.float() + BN_EPS)      # = (x - running_mean) / sqrt(running_var + eps)
        bias = -bn.running_mean.float() * scale
    else:
        scale = (bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + BN_EPS))
        bias = bn.bias.detach().float() - bn.running_mean.float() * scale
    if conv_bias is not None:
        bias = bias + conv_bias.detach().float() * scale
    return scale.contiguous(), bias.contiguous()


def _interleave(v, g):
    """bias permutation matching essb_pack_weight(interleave=g): packed[ch*g + j] = v[j*G + ch]."""
    G = v.numel() // g
    return v.detach().float().view(g, G).t().contiguous().view(-1)


class E2VIDRecurrent(nn.Module):
    """Recurrent U-Net event encoder / image reconstructor (E2VID), B200-native."""

    def __init__(self, config, mode=None):
        super().__init__()
        self.config = config
        assert 'num_bins' in config
        self.num_bins = int(config['num_bins'])                                   # model.py:13-14
        self.skip_type = str(config.get('skip_type', 'sum'))
        self.num_encoders = int(config.get('num_encoders', 4))
        self.base_num_channels = int(config.get('base_num_channels', 32))
        self.num_residual_blocks = int(config.get('num_residual_blocks', 2))
        self.norm = str(config['norm']) if 'norm' in config else None
        self.use_upsample_conv = bool(config.get('use_upsample_conv', True))
        self.recurrent_block_type = str(config.get('recurrent_block_type', 'convlstm'))
        if self.norm not in (None, 'BN', 'IN'):
            raise ValueError("E2VIDRecurrent: norm=%r (None, 'BN' or 'IN', submodules.py:19-22)" % (self.norm,))
        if self.recurrent_block_type not in ('convlstm', 'convgru'):
            raise ValueError(self.recurrent_block_type)
        if self.num_encoders < 3:
            raise ValueError('UNetRecurrent.forward needs >= 3 encoders (latent dict, unet.py:172)')
        if self.num_residual_blocks < 1:
            raise NotImplementedError('num_residual_blocks == 0')
        self.unetrecurrent = _UNetRecurrent(dict(
            num_bins=self.num_bins, skip_type=self.skip_type, num_encoders=self.num_encoders,
            base_num_channels=self.base_num_channels, num_residual_blocks=self.num_residual_blocks, norm=self.norm,
            use_upsample_conv=self.use_upsample_conv, recurrent_block_type=self.recurrent_block_type))
        self.mode = mode or default_mode()
        self._packed = None
        self._packed_key = None
        self._planes = {}   # data_ptr of a returned hidden state -> (tensor, hi, lo) for the next window

    # ------------------------------------------------------------------------------ weight packing
    def _key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers())) + (self.mode, self._enc0_hf8(), os.environ.get('ESS_B200_MERGE_PHASES', '1'))

    def _pack(self):
        key = self._key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        u = self.unetrecurrent
        dev = u.head.conv2d.weight.device
        P = {}
        cpad = (self.num_bins + 7) // 8 * 8
        wh = torch.zeros((self.base_num_channels, cpad, 5, 5), device=dev)
        wh[:, :self.num_bins] = u.head.conv2d.weight.detach().float()
        P['head'] = (ops.pack_weight(wh), u.head.conv2d.bias.detach().float().contiguous(), cpad)
        tc = self.mode != 'fp32'
        hf8 = self.mode == 'f16f8'

        def pk(w, scale=None, **kw):
            """(hi, lo, KinP, acc_scale) in the mode's operand format"""
            if hf8:
                return ops.pack_weight_tc_hf8(w, scale, **kw)
            return ops.pack_weight_tc(w, scale, **kw) + (0.0,)
        P['head_tc'] = None
        if tc and cpad in (8, 16) and self.base_num_channels % 32 == 0 and self.base_num_channels <= 256:
            # per kernel row ky the head conv reads the 8-pixel x cpad-channel window starting at x-2 as one
            # contiguous K-chunk (ops.window_view): virtual weight [Cout][8*cpad][ky], zero for kx >= 5
            # `grp` adjacent output pixels share one window (output i uses window pixels i .. i+4) and are
            # computed as grp*Cout accumulator columns: tcgen05.mma costs the same for N = 32 and N = 128,
            # so widening N to 128 does 4 pixels per instruction.  [N,H,W,base] == [N,H,W/grp,grp*base].
            base = self.base_num_channels
            grp = 4 if 4 * base <= 256 else (2 if 2 * base <= 256 else 1)
            wk = u.head.conv2d.weight.detach().float().permute(0, 3, 1, 2)  # [co,kx,ci,ky]
            wv = torch.zeros((grp, base, 8, cpad, 5), device=dev)
            for i in range(grp):
                wv[i, :, i:i + 5, :self.num_bins] = wk
            wv = wv.reshape(grp * base, 8 * cpad, 5, 1).contiguous()
            hi, lo, kinp = ops.pack_weight_tc(wv)
            P['head_tc'] = dict(hi=hi, lo=lo, k_per_tap=kinp, cpad=cpad, grp=grp,
                                bias=P['head'][1].repeat(grp).contiguous())
        for i, enc in enumerate(u.encoders):
            bn = getattr(enc.conv, 'norm_layer', None)
            scale, bias = _bn_fold(enc.conv.conv2d.bias, bn, None, dev)
            e = {'bias': bias, 'w': ops.pack_weight(enc.conv.conv2d.weight, scale)}
            cin, cout = enc.conv.conv2d.in_channels, enc.conv.conv2d.out_channels
            e['tc'] = None
            if tc and cout % 64 == 0 and (cin % 64 == 0 or cin == 32):
                w = enc.conv.conv2d.weight.detach().float()
                fold = cin == 32
                if fold:   # horizontal pixel pairs become one 64-channel super-pixel (see ops.parity_view)
                    w6 = torch.zeros((cout, cin, 5, 6), device=dev)
                    w6[..., :5] = w
                    w = w6.view(cout, cin, 5, 3, 2).permute(0, 4, 1, 2, 3).reshape(cout, 2 * cin, 5, 3).contiguous()
                if i == 0 and hf8 and not self._enc0_hf8():
                    # f16f8 mode, first encoder conv in bf16x3: its input planes come from the head conv, whose epilogue
                    # is the head's bottleneck -- the cheaper bf16 hi/lo split wins there, and this N = 64 layer is bound
                    # by the shared-memory operand feed, not by its MMA count (measured, profiles/r02_*)
                    hi, lo, kinp = ops.pack_weight_tc(w, scale)
                    e['tc'] = dict(hi=hi, lo=lo, k_per_tap=kinp, fold=fold, T=w.shape[2] * w.shape[3], sc=0.0, passes=3)
                else:
                    hi, lo, kinp, sc = pk(w, scale)
                    e['tc'] = dict(hi=hi, lo=lo, k_per_tap=kinp, fold=fold, T=w.shape[2] * w.shape[3], sc=sc)
            rb = enc.recurrent_block
            C = cout
            if self.recurrent_block_type == 'convlstm':
                W = rb.Gates.weight
                e['lstm_w'] = ops.pack_weight(W, interleave=4)
                e['lstm_w_x'] = ops.pack_weight(W[:, :C].contiguous(), interleave=4)
                e['lstm_b'] = _interleave(rb.Gates.bias, 4)
                e['lstm_tc'] = None
                if tc and C % 64 == 0:
                    hi, lo, kinp, sc = pk(W, interleave=4)
                    e['lstm_tc'] = dict(hi=hi, lo=lo, k_per_tap=kinp, sc=sc)
            else:
                wur = torch.cat([rb.update_gate.weight, rb.reset_gate.weight], 0).detach()
                bur = torch.cat([rb.update_gate.bias, rb.reset_gate.bias], 0).detach()
                e['gru_ur_w'] = ops.pack_weight(wur, interleave=2)
                e['gru_ur_w_x'] = ops.pack_weight(wur[:, :C].contiguous(), interleave=2)
                e['gru_ur_b'] = _interleave(bur, 2)
                e['gru_o_w'] = ops.pack_weight(rb.out_gate.weight)
                e['gru_o_w_x'] = ops.pack_weight(rb.out_gate.weight[:, :C].contiguous())
                e['gru_o_b'] = rb.out_gate.bias.detach().float().contiguous()
                e['gru_tc'] = None
                if tc and C % 64 == 0:
                    uh, ul, kinp, usc = pk(wur, interleave=2)
                    oh_, ol_, _, osc = pk(rb.out_gate.weight)
                    e['gru_tc'] = dict(ur_hi=uh, ur_lo=ul, o_hi=oh_, o_lo=ol_, k_per_tap=kinp, ur_sc=usc, o_sc=osc)
            P['enc%d' % i] = e
        for j, rbk in enumerate(u.resblocks):
            s1, b1 = _bn_fold(rbk.conv1.bias, getattr(rbk, 'bn1', None), None, dev)
            s2, b2 = _bn_fold(rbk.conv2.bias, getattr(rbk, 'bn2', None), None, dev)
            P['res%d' % j] = (ops.pack_weight(rbk.conv1.weight, s1), b1, ops.pack_weight(rbk.conv2.weight, s2), b2)
            if tc and self.norm != 'IN':      # norm='IN': the two resblocks normalise per sample -> fp32 image decoder
                P['res%d_tc' % j] = pk(rbk.conv1.weight, s1) + pk(rbk.conv2.weight, s2)
        for i, dec in enumerate(u.decoders):
            bn = getattr(dec, 'norm_layer', None)
            if self.use_upsample_conv:
                scale, bias = _bn_fold(dec.conv2d.bias, bn, None, dev)
                P['dec%d' % i] = (ops.pack_weight(dec.conv2d.weight, scale), bias, dec.conv2d.out_channels)
                if tc and dec.conv2d.in_channels % 64 == 0 and dec.conv2d.out_channels % 32 == 0:
                    P['dec%d_tc' % i] = pk(dec.conv2d.weight, scale)
            else:
                scale, bias = _bn_fold(dec.transposed_conv2d.bias, bn, None, dev)
                P['dec%d' % i] = (ops.pack_weight(dec.transposed_conv2d.weight, scale, transposed_layout=True), bias,
                                  dec.transposed_conv2d.out_channels)
                if tc and dec.transposed_conv2d.in_channels % 64 == 0 and dec.transposed_conv2d.out_channels % 32 == 0:
                    P['dec%d_tc' % i] = pk(dec.transposed_conv2d.weight, scale, transposed_layout=True)
                    if os.environ.get('ESS_B200_MERGE_PHASES', '1') != '0' and 4 * dec.transposed_conv2d.out_channels <= 1024:
                        # the four sub-pixel phases as ONE launch of N = 4 * Cout (essb_conv_tc.phase_cout)
                        wm = ops.merge_convT_phases(dec.transposed_conv2d.weight)
                        P['dec%d_tcm' % i] = pk(wm, scale.repeat(4) if scale is not None else None) + \
                            (bias.repeat(4).contiguous() if bias is not None else None,)
        scale, bias = _bn_fold(u.pred.conv2d.bias, getattr(u.pred, 'norm_layer', None), None, dev)
        P['pred'] = (ops.pack_weight(u.pred.conv2d.weight, scale), bias)
        wp = u.pred.conv2d.weight.detach().float().reshape(1, -1)
        if scale is not None:
            wp = wp * scale.view(1, 1)
        P['pred_pw'] = wp.contiguous() if wp.shape[1] in (32, 64) else None    # streaming 1x1 kernel (pw_conv.cu)
        self._packed, self._packed_key = P, key
        return P

    # ------------------------------------------------------------------------------ row-stacked levels
    @staticmethod
    def _stack_plan(N, H, W, ne):
        """Per encoder level i (output oh x ow): rows per image `hp` of the level's buffers (hp > oh = row-stacked, see
        essb_conv_tc.row_period) and whether the level's stride-2 conv runs on the tall view as well.  The tensor-core
        kernels tile the output in 16-row patches PER IMAGE: a 55-row map (DSEC at 1/8 scale) pads to 64 rows (14 %), a
        25-row map (DDD17) to 32.  Stacking the batch vertically with shared zero rows between the images (8 x 56 = 448 =
        28 patches instead of 32) removes the rounding; the zero rows are the convolution's padding for both neighbours.
        A level is stacked when that saves patches for its ConvLSTM, or lets the NEXT level's conv run tall
        (its input period must be exactly twice its output period).  ESS_B200_STACK=0 disables it."""
        if os.environ.get('ESS_B200_STACK', '1') == '0' or N < 2:
            return [(H >> (i + 1), False) for i in range(ne)]
        ohs = [H >> (i + 1) for i in range(ne)]
        hp_last = ohs[-1] + 1
        hps = [hp_last << (ne - 1 - i) for i in range(ne)]

        def saves(i):
            return -(-N * hps[i] // 16) < N * -(-ohs[i] // 16)
        plan = []
        for i in range(ne):
            own = saves(i)
            for_next = i + 1 < ne and saves(i + 1)
            plan.append(hps[i] if (own or for_next) else ohs[i])
        out = []
        for i in range(ne):
            stacked = plan[i] > ohs[i]
            tall_conv = stacked and i > 0 and plan[i - 1] == 2 * plan[i] and plan[i - 1] > ohs[i - 1] and saves(i)
            out.append((plan[i], tall_conv))
        return out

    def _zero_planes(self, key, shape, device, avoid=None):
        """Persistent zero-initialised operand-plane pair (hi, lo) for a row-stacked level: its zero rows are never
        written (masked stores), so it is cleared once.  Two alternating pairs per key; `avoid` = data_ptr of the pair
        being READ by the launch that will write the returned one (the previous hidden state)."""
        cache = self.__dict__.setdefault('_stack_bufs', {})
        k = (key, tuple(shape), str(device))
        ent = cache.get(k)
        if ent is None:
            ent = [tuple(torch.zeros(shape, device=device, dtype=torch.bfloat16) for _ in range(2)) for _ in range(2)]
            cache[k] = ent
        return ent[1] if ent[0][0].data_ptr() == avoid else ent[0]

    # --------------------------------------------------------------------------------- state helpers
    @staticmethod
    def _to_nhwc(t):
        v = t.permute(0, 2, 3, 1)
        return v if v.is_contiguous() else v.contiguous()

    def _hidden_planes(self, h_nhwc):
        ent = self._planes.get(h_nhwc.data_ptr())
        if ent is not None and len(ent) == 4 and ent[0].data_ptr() == h_nhwc.data_ptr() and ent[0]._version == ent[3] \
                and ent[0].shape == h_nhwc.shape and h_nhwc.is_contiguous():
            return ent[1], ent[2]
        N, H, W, _ = h_nhwc.shape
        return ops.split_bf16(Seg(h_nhwc), N, H, W, fmt=self._fmt())

    @staticmethod
    def _enc0_hf8():
        """f16f8 mode: does the first encoder conv consume hf8 planes (ESS_B200_F16F8_ENC0=hf8) or bf16 hi/lo planes
        in three passes (default)?"""
        return os.environ.get('ESS_B200_F16F8_ENC0', 'bf16x3') == 'hf8'

    def _fmt_enc0(self):
        """operand-plane format of the head conv's output (= the first encoder conv's input)"""
        return ops.PLANES_HF8 if (self.mode == 'f16f8' and self._enc0_hf8()) else ops.PLANES_BF16

    def _fmt(self):
        """operand-plane format of the activations that travel between the tcgen05 launches of this network"""
        return ops.PLANES_HF8 if self.mode == 'f16f8' else ops.PLANES_BF16

    # --------------------------------------------------------------------------------------- forward
    def forward(self, event_tensor, prev_states, with_image=True):
        """event_tensor [N, num_bins, H, W] (H, W multiples of 2^num_encoders), prev_states None or a
        list of (hidden, cell) tuples (ConvLSTM) / tensors (ConvGRU).  Returns (img [N,1,H,W] or None
        when with_image=False, states, latent {1,2,4,8}) exactly as unet.py:145-181."""
        ops.require_cuda_any(event_tensor)
        N, Cb, H, W = event_tensor.shape
        if Cb != self.num_bins:
            raise RuntimeError('expected %d input channels, got %d' % (self.num_bins, Cb))
        f = 2 ** self.num_encoders
        if H % f or W % f:
            raise RuntimeError('H, W must be multiples of %d (CropParameters pads to this)' % f)
        if torch.is_grad_enabled() and self._wants_grad(event_tensor, prev_states):
            return self._forward_bptt(event_tensor, prev_states, with_image)
        with torch.no_grad(), ops.on_device_of(event_tensor):
            buf = self.head_planes_buffer(N, H, W, event_tensor.device)
            if buf is not None:      # tensor-core head: convert straight into its operand format
                ev = event_tensor.float().contiguous()
                ops.event_prepare_planes(ev, None, False, H, W, 0, 0, buf)
                return self.forward_planes(buf, H, W, prev_states, with_image)
            cpad = (self.num_bins + 7) // 8 * 8
            x = ops.nchw_to_nhwc(event_tensor, cpad)
            return self.forward_nhwc(x, prev_states, with_image)

    # ------------------------------------------------------- differentiable path (back-propagation through time)
    def _wants_grad(self, event_tensor, prev_states):
        if event_tensor.requires_grad or any(p.requires_grad for p in self.parameters()):
            return True
        for st in (prev_states or []):
            for t in (st if isinstance(st, (tuple, list)) else (st,)):
                if t is not None and t.requires_grad:
                    return True
        return False

    def _forward_bptt(self, event_tensor, prev_states, with_image=True):
        """forward() with autograd: gradients w.r.t. the parameters, the incoming states and the event tensor, so the
        encoder can be trained THROUGH the recurrence like the reference module (SURVEY.md s8f "later": no reference
        trainer does it -- they freeze the encoder and call it under no_grad, which takes the fused path).  Every
        convolution runs on this library's kernels with its hand-written input / weight gradients (`style_encoder._ConvFn`:
        tcgen05 for the 3x3 gate convolutions, in <= 256-output-channel slices, CUDA cores for the 5x5 ones); the gate
        non-linearities, the eval-mode BatchNorm affine and the state update are autograd-recorded elementwise ops, and the
        4C-channel gate tensor IS materialised (the backward needs it) -- this is the slow, memory-hungry path by design.
        `img` is computed without gradient.  ConvLSTM and ConvGRU; BatchNorm always uses its running statistics."""
        import torch.nn.functional as F
        from .style_encoder import _ConvFn
        if any(isinstance(m_, nn.BatchNorm2d) and m_.training for m_ in self.modules()):
            _warn_once('E2VIDRecurrent: BatchNorm layers are in training mode but this implementation always normalises with '
                       'the running statistics (the reference trainers call .eval(), ess_supervised_trainer.py:47)')
        _warn_once('E2VIDRecurrent: gradient mode is on and the encoder has trainable parameters / differentiable inputs -- '
                   'taking the differentiable (BPTT) path, several times slower than the fused inference path; wrap the '
                   'call in torch.no_grad() or freeze the parameters when no encoder gradient is needed')
        mode = 'bf16x3' if self.mode == 'f16f8' else self.mode
        u = self.unetrecurrent
        ne = self.num_encoders
        N, Cb, H, W = event_tensor.shape
        prev_states = list(prev_states) if prev_states is not None else [None] * ne
        with ops.on_device_of(event_tensor):
            cpad = (Cb + 3) // 4 * 4
            x = F.pad(event_tensor.float().permute(0, 2, 3, 1), (0, cpad - Cb)).contiguous()
            wh = F.pad(u.head.conv2d.weight, (0, 0, 0, 0, 0, cpad - Cb))
            head = torch.relu(_ConvFn.apply(x, wh, 1, 2, mode) + u.head.conv2d.bias)              # unet.py:131-132,153
            cur = head
            blocks, states = [], []
            for i, enc in enumerate(u.encoders):
                conv = enc.conv.conv2d
                y = _ConvFn.apply(cur.contiguous(), conv.weight, 2, 2, mode)                      # submodules.py:7-31
                if conv.bias is not None:
                    y = y + conv.bias
                bn = getattr(enc.conv, 'norm_layer', None)
                if isinstance(bn, nn.InstanceNorm2d):                                              # norm='IN': running statistics
                    y = (y - bn.running_mean) / torch.sqrt(bn.running_var + BN_EPS)
                elif bn is not None:                                                               # eval-mode BatchNorm2d
                    y = (y - bn.running_mean) * (bn.weight / torch.sqrt(bn.running_var + BN_EPS)) + bn.bias
                y = torch.relu(y)
                C = conv.out_channels
                st = prev_states[i]
                if self.recurrent_block_type == 'convgru':                                         # submodules.py:255-273
                    rb = enc.recurrent_block
                    h_prev = torch.zeros_like(y) if st is None else st.float().permute(0, 2, 3, 1).contiguous()
                    xh = torch.cat([y, h_prev], -1)
                    upd = torch.sigmoid(_ConvFn.apply(xh, rb.update_gate.weight, 1, 1, mode) + rb.update_gate.bias)
                    rst = torch.sigmoid(_ConvFn.apply(xh, rb.reset_gate.weight, 1, 1, mode) + rb.reset_gate.bias)
                    out = torch.tanh(_ConvFn.apply(torch.cat([y, h_prev * rst], -1), rb.out_gate.weight, 1, 1, mode) +
                                     rb.out_gate.bias)
                    h = h_prev * (1 - upd) + out * upd
                    blocks.append(h)
                    states.append(ops.as_nchw(h))
                    cur = h
                    continue
                if st is None:                                                                     # submodules.py:196-207
                    h_prev = torch.zeros_like(y)
                    c_prev = torch.zeros_like(y)
                else:
                    h_prev = st[0].float().permute(0, 2, 3, 1).contiguous()
                    c_prev = st[1].float().permute(0, 2, 3, 1).contiguous()
                xh = torch.cat([y, h_prev], -1)                                                    # :212
                gw, gb = enc.recurrent_block.Gates.weight, enc.recurrent_block.Gates.bias
                parts = [_ConvFn.apply(xh, gw[j:j + 256], 1, 1, mode) for j in range(0, 4 * C, 256)]
                gates = (parts[0] if len(parts) == 1 else torch.cat(parts, -1)) + gb               # :213
                g_in, g_rem, g_out, g_cell = gates.chunk(4, -1)                                    # :216
                c = torch.sigmoid(g_rem) * c_prev + torch.sigmoid(g_in) * torch.tanh(g_cell)       # :219-227
                h = torch.sigmoid(g_out) * torch.tanh(c)                                           # :228
                blocks.append(h)
                states.append((ops.as_nchw(h), ops.as_nchw(c)))
                cur = h
            latent = {1: ops.as_nchw(head), 2: ops.as_nchw(blocks[0]), 4: ops.as_nchw(blocks[1]), 8: ops.as_nchw(blocks[2])}
            img = None
            if with_image:
                with torch.no_grad():
                    P = self._pack()
                    img = ops.as_nchw(self._image_decoder(P, head.detach().contiguous(), [b.detach().contiguous() for b in blocks],
                                                          N, H >> ne, W >> ne))
        return img, states, latent

    def head_planes_buffer(self, N, H, W, device):
        """Cached zero-bordered bf16 hi/lo input buffer of the tensor-core head conv for an [N, *, H, W]
        input (see ops.head_planes_alloc), or None when the head runs on the fp32 kernel."""
        P = self._pack()
        if P['head_tc'] is None:
            return None
        key = (N, H, W, str(device))
        buf = self._in_planes.get(key) if hasattr(self, '_in_planes') else None
        if buf is None:
            buf = ops.head_planes_alloc(N, H, W, P['head_tc']['cpad'], device)
            self._in_planes = {key: buf}
        return buf

    def forward_planes(self, in_planes, H, W, prev_states, with_image=True, want_head=True):
        """forward() for an input already in the head conv's operand format (essb_event_prepare_planes).
        want_head=False (only with with_image=False): the fp32 head activation -- latent[1], needed by nobody
        between the windows of an unroll -- is not written (latent[1] is None); the head conv then only emits
        the bf16 operand planes of the first encoder."""
        return self._forward_impl(None, in_planes, in_planes[0].shape[0], H, W, prev_states, with_image,
                                  want_head=want_head or with_image)

    def forward_nhwc(self, x, prev_states, with_image=True):
        """Same as forward() for an already pixel-major input [N, H, W, ceil8(num_bins)] (zero-padded
        channels), as produced by the fused event pre-processing kernel."""
        return self._forward_impl(x, None, x.shape[0], x.shape[1], x.shape[2], prev_states, with_image)

    def _head_tc(self, P, in_planes, N, H, W, head, planes, passes):
        """head conv5x5 + bias + ReLU (unet.py:131-132,153) on the tcgen05 kernel: 5 taps (kernel rows), each
        one 8*cpad-element K-chunk read through the overlapping-stride window view."""
        tcw = P['head_tc']
        grp = tcw['grp']
        base, W = self.base_num_channels * grp, W // grp
        d = ConvTc()
        ops.window_view(d.views[0], in_planes[0], in_planes[1], H, W * grp, grp)
        d.n_views, d.nseg = 1, 1
        d.seg_C[0], d.seg_view0[0], d.seg_koff[0] = 8 * tcw['cpad'], 0, 0
        d.k_per_tap, d.n_w_taps, d.w_rows = tcw['k_per_tap'], 5, tcw['hi'].shape[0]
        d.w_hi, d.w_lo, d.bias = ops._p(tcw['hi']), ops._p(tcw['lo']), ops._p(tcw['bias'])
        d.out, d.ldo = ops._p(head), base
        d.out_hi, d.out_lo, d.ld_planes = ops._p(planes[0]), ops._p(planes[1]), base
        d.N, d.OH, d.OW, d.Cout = N, H, W, base
        d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = H, W, 1, 0, 1, 0
        # f16f8 mode: the head itself runs bf16x3 (its input planes are bf16 hi/lo) but emits hf8 planes for encoder 0
        d.epilogue, d.act, d.passes, d.bw_log2 = EPI_LINEAR, ACT_RELU, (3 if passes == 2 else passes), ops.pick_bw_log2(W, H)
        d.planes_fmt = self._fmt_enc0()
        d.ntaps = 5
        for ky in range(5):
            d.dy[ky], d.dx[ky], d.view[ky], d.widx[ky] = ky, 0, 0, ky
        ops.conv_tc(d, tag='head_tc', device=head.device if head is not None else planes[0].device)

    def _forward_impl(self, x, in_planes, N, H, W, prev_states, with_image, want_head=True):
        with ops.on_device_of(x if x is not None else in_planes[0]):
            return self._forward_on_device(x, in_planes, N, H, W, prev_states, with_image, want_head)

    def _forward_on_device(self, x, in_planes, N, H, W, prev_states, with_image, want_head=True):
        P = self._pack()
        u = self.unetrecurrent
        ne = self.num_encoders
        lstm = self.recurrent_block_type == 'convlstm'
        tc_mode = self.mode != 'fp32'
        passes = ops.PASSES.get(self.mode, 1)
        if prev_states is None:
            prev_states = [None] * ne
        base = self.base_num_channels
        dev = x.device if x is not None else in_planes[0].device

        # head: conv5x5 + bias + ReLU (unet.py:131-132,153)
        wh, bh, cpad = P['head']
        want_planes = tc_mode and P['enc0']['tc'] is not None
        planes = None
        if want_planes or in_planes is not None:
            planes = (torch.empty((N, H, W, base), device=dev, dtype=torch.bfloat16),
                      torch.empty((N, H, W, base), device=dev, dtype=torch.bfloat16))
        if in_planes is not None:
            head = torch.empty((N, H, W, base), device=dev, dtype=torch.float32) if (want_head or not want_planes) else None
            self._head_tc(P, in_planes, N, H, W, head, planes, passes)
        else:
            hf8 = self._fmt_enc0() == ops.PLANES_HF8  # the fp32 kernel's epilogue only writes bf16 hi/lo planes
            head, _, _, _ = ops.conv([Seg(x, C=cpad)], wh, bh, N, H, W, H, W, base, ops.taps_conv(5, 2), act=ACT_RELU,
                                     planes=None if hf8 else planes)
            if hf8 and planes is not None:
                g = max(1, 64 // base)                # hf8 rows are 64 channels: fold g pixels (W % 8 == 0)
                shp = (N, H, W // g, base * g)
                ops.split_bf16(Seg(head.view(shp)), N, H, W // g, planes[0].view(shp), planes[1].view(shp), fmt=ops.PLANES_HF8)

        blocks, states, block_planes = [], [], []
        plan = self._stack_plan(N, H, W, ne) if (tc_mode and lstm) else [(H >> (i + 1), False) for i in range(ne)]
        prev_rows = H
        cur, cur_planes = head, planes
        new_planes = {}
        h_in, w_in = H, W
        for i in range(ne):
            e = P['enc%d' % i]
            cin, cout = base * 2 ** i, base * 2 ** (i + 1)
            oh, ow = h_in // 2, w_in // 2
            st = prev_states[i]
            use_tc = tc_mode and e['tc'] is not None and cur_planes is not None and \
                (e['lstm_tc'] if lstm else e['gru_tc']) is not None
            if use_tc and not lstm:
                xh = torch.empty((N, oh, ow, cout), device=dev, dtype=torch.bfloat16)
                xl = torch.empty_like(xh)
                self._enc_conv_tc(e, cur_planes, N, h_in, w_in, cout, xh, xl, passes)
                hp = self._to_nhwc(st) if st is not None else None
                hp_planes = self._hidden_planes(hp) if hp is not None else None
                h, hh, hl = self._gru_tc(e, (xh, xl), hp, hp_planes, N, oh, ow, cout, passes)
                new_planes[h.data_ptr()] = (h, hh, hl, h._version)
                cur, cur_planes = h, (hh, hl)
                state = ops.as_nchw(h)
            elif use_tc and plan[i][0] > oh:
                # row-stacked level (see _stack_plan): buffers [N, rows, ow, C] with rows - oh zero rows per image
                rows, tall_conv = plan[i]
                xh, xl = self._zero_planes(('x', i), (N, rows, ow, cout), dev)
                self._enc_conv_tc(e, cur_planes, N, h_in, w_in, cout, xh, xl, passes, rows_out=rows,
                                  tall=(prev_rows if tall_conv else 0))
                h_full = c_full = hp_planes = None
                if st is not None:
                    ent = self._planes.get(st[0].data_ptr())
                    if ent is not None and len(ent) == 6 and ent[5] == rows and ent[0]._version == ent[3] \
                            and tuple(st[0].shape) == (N, cout, oh, ow) and st[1].data_ptr() == ent[4].data_ptr():
                        hp_planes, c_full = (ent[1], ent[2]), ent[4]          # our own previous output: already stacked
                    else:                                                      # foreign states: stack them once
                        h_full = torch.zeros((N, rows, ow, cout), device=dev, dtype=torch.float32)
                        c_full = torch.zeros_like(h_full)
                        h_full[:, :oh] = self._to_nhwc(st[0])
                        c_full[:, :oh] = self._to_nhwc(st[1])
                        hp_planes = ops.split_bf16(Seg(h_full.view(1, N * rows, ow, cout)), 1, N * rows, ow, fmt=self._fmt())
                        hp_planes = tuple(t.view(N, rows, ow, cout) for t in hp_planes)
                out_planes = self._zero_planes(('h', i), (N, rows, ow, cout), dev,
                                               avoid=hp_planes[0].data_ptr() if hp_planes is not None else None)
                h_full, c_full, hh, hl = self._lstm_tc(e, (xh, xl), hp_planes, c_full, N, oh, ow, cout, passes,
                                                       rows=rows, out_planes=out_planes)
                h, c = h_full[:, :oh], c_full[:, :oh]
                new_planes[h.data_ptr()] = (h, hh, hl, h_full._version, c_full, rows)
                cur, cur_planes = h, (hh[:, :oh], hl[:, :oh])
                state = (ops.as_nchw(h), ops.as_nchw(c))
            elif use_tc:
                xh = torch.empty((N, oh, ow, cout), device=dev, dtype=torch.bfloat16)
                xl = torch.empty_like(xh)
                self._enc_conv_tc(e, cur_planes, N, h_in, w_in, cout, xh, xl, passes)
                hp = cp = None
                hp_planes = None
                if st is not None:
                    hp, cp = self._to_nhwc(st[0]), self._to_nhwc(st[1])
                    hp_planes = self._hidden_planes(hp)
                h, c, hh, hl = self._lstm_tc(e, (xh, xl), hp_planes, cp, N, oh, ow, cout, passes)
                new_planes[h.data_ptr()] = (h, hh, hl, h._version)
                cur, cur_planes = h, (hh, hl)
                state = (ops.as_nchw(h), ops.as_nchw(c))
            else:
                if cur is None:
                    raise RuntimeError('internal: fp32 activation missing')
                if not cur.is_contiguous():
                    cur = cur.contiguous()
                if tc_mode:
                    _warn_once('E2VIDRecurrent(mode=%r): encoder level %d (%d -> %d channels) runs on the fp32 CUDA-core '
                               'kernels (~10x slower; the tensor-core path needs base_num_channels %% 32 == 0 and '
                               'num_bins <= 16)' % (self.mode, i, cin, cout))
                xi, _, _, _ = ops.conv([Seg(cur)], e['w'], e['bias'], N, h_in, w_in, oh, ow, cout, ops.taps_conv(5, 2),
                                       stride=2, act=ACT_RELU)
                t3 = ops.taps_conv(3, 1)
                if lstm:
                    hp = cp = None
                    if st is not None:
                        hp, cp = self._to_nhwc(st[0]), self._to_nhwc(st[1])
                    segs = [Seg(xi)] + ([Seg(hp)] if hp is not None else [])
                    h, c, _, _ = ops.conv(segs, e['lstm_w'] if hp is not None else e['lstm_w_x'], e['lstm_b'], N, oh, ow,
                                          oh, ow, 4 * cout, t3, epilogue=EPI_LSTM, aux0=cp)
                    state = (ops.as_nchw(h), ops.as_nchw(c))
                else:
                    hp = self._to_nhwc(st) if st is not None else None
                    segs = [Seg(xi)] + ([Seg(hp)] if hp is not None else [])
                    upd, hr, _, _ = ops.conv(segs, e['gru_ur_w'] if hp is not None else e['gru_ur_w_x'], e['gru_ur_b'],
                                             N, oh, ow, oh, ow, 2 * cout, t3, epilogue=EPI_GRU_UR, aux0=hp)
                    segs = [Seg(xi)] + ([Seg(hr)] if hp is not None else [])
                    h, _, _, _ = ops.conv(segs, e['gru_o_w'] if hp is not None else e['gru_o_w_x'], e['gru_o_b'], N, oh,
                                          ow, oh, ow, cout, t3, epilogue=EPI_GRU_OUT, aux0=hp, aux1=upd)
                    state = ops.as_nchw(h)
                cur, cur_planes = h, None
            blocks.append(cur)
            if cur_planes is not None:
                block_planes.append(cur_planes)
            states.append(state)
            prev_rows = plan[i][0] if (use_tc and lstm) else oh
            h_in, w_in = oh, ow
        self._planes = new_planes

        latent = {1: ops.as_nchw(head) if head is not None else None, 2: ops.as_nchw(blocks[0]), 4: ops.as_nchw(blocks[1]),
                  8: ops.as_nchw(blocks[2])}                                       # unet.py:172
        if not with_image:
            return None, states, latent
        cmax = base * 2 ** ne
        if tc_mode and len(block_planes) == ne and base % 32 == 0 and cmax <= 256 * 4 and base * 2 % 64 == 0 \
                and all('dec%d_tc' % i in P for i in range(ne)) and 'res0_tc' in P:
            img = self._image_decoder_tc(P, head, blocks, block_planes, N, h_in, w_in, passes)
        else:
            if tc_mode:
                _warn_once('E2VIDRecurrent(mode=%r): this configuration (base_num_channels=%d, norm=%r) runs its image decoder '
                           'on the fp32 CUDA-core kernels (~10x slower than the tensor-core path; that one needs channel '
                           'counts that are multiples of 32 with 64-wide decoder inputs and norm None / BN)'
                           % (self.mode, base, self.norm))
            img = self._image_decoder(P, head, blocks, N, h_in, w_in)
        return ops.as_nchw(img), states, latent

    # --------------------------------------------------------- image decoder on the tcgen05 kernel
    def _image_decoder_tc(self, P, head, blocks, block_planes, N, h, w, passes):
        """2 ResidualBlocks + 3 up-sampling layers on the tensor-core kernel; activations travel between layers as
        operand planes written by the epilogues.  TransposedConvLayer (submodules.py:34-62) = 4 sub-pixel phases with
        strided epilogue stores; UpsampleConvLayer (submodules.py:65-93) = bilinear x2 (HBM-bound kernel) -> planes ->
        25-tap conv.  skip_type 'sum' (unet.py:175,179): epilogue adds; 'concat': the skip tensor's planes are a second
        K segment of the same launch (no torch.cat).  The 1x1 prediction conv stays on the streaming / fp32 kernel."""
        ne, base = self.num_encoders, self.base_num_channels
        cmax = base * 2 ** ne
        dev = blocks[-1].device
        t3 = ops.taps_conv(3, 1)
        fmt = self._fmt()
        concat = self.skip_type != 'sum'
        up = self.use_upsample_conv

        def new_planes(hh, ww, c):
            return (torch.empty((N, hh, ww, c), device=dev, dtype=torch.bfloat16),
                    torch.empty((N, hh, ww, c), device=dev, dtype=torch.bfloat16))

        blocks = [b if b.is_contiguous() else b.contiguous() for b in blocks]   # residual / skip sources must be dense
        x, xp = blocks[-1], block_planes[-1]
        nres = self.num_residual_blocks
        for j in range(nres):
            hi1, lo1, k1, sc1, hi2, lo2, k2, sc2 = P['res%d_tc' % j]
            b1, b2 = P['res%d' % j][1], P['res%d' % j][3]
            tp = new_planes(h, w, cmax)
            ops.conv_tc_dense(xp, hi1, lo1, k1, t3, N, h, w, cmax, passes, bias=b1, act=ACT_RELU, want_out=False,
                              out_planes=tp, tag='img_tc', acc_scale=sc1, planes_fmt=fmt)
            post = blocks[ne - 1] if (j == nres - 1 and not concat) else None
            np_ = new_planes(h, w, cmax)
            x = ops.conv_tc_dense(tp, hi2, lo2, k2, t3, N, h, w, cmax, passes, bias=b2, act=ACT_RELU, res_pre=x,
                                  res_post=post, out_planes=np_, tag='img_tc', acc_scale=sc2, planes_fmt=fmt)
            xp = np_
        for i in range(ne):
            hi, lo, k, sc = P['dec%d_tc' % i]
            bd, cout = P['dec%d' % i][1], P['dec%d' % i][2]
            skip_next = blocks[ne - i - 2] if i < ne - 1 else head
            post = None if concat else skip_next
            if up:
                srcs = [x] + ([blocks[ne - i - 1]] if concat else [])
                segs = [ops.split_bf16(Seg(ops.bilinear_up2(s_)), N, 2 * h, 2 * w, fmt=fmt) for s_ in srcs]
                x = ops.conv_tc_dense(segs, hi, lo, k, ops.taps_conv(5, 2), N, 2 * h, 2 * w, cout, passes, bias=bd,
                                      act=ACT_RELU, res_post=post, tag='img_tc', acc_scale=sc, planes_fmt=fmt)
                xp = None
            else:
                segs = [xp] + ([block_planes[ne - i - 1]] if concat else [])
                out = torch.empty((N, 2 * h, 2 * w, cout), device=dev, dtype=torch.float32)
                op = new_planes(2 * h, 2 * w, cout) if i < ne - 1 else None
                merged = P.get('dec%d_tcm' % i)
                if merged is not None:
                    mhi, mlo, mk, msc, mb = merged
                    cin_t = sum(sp[0].shape[-1] for sp in segs)
                    ops.conv_tc_dense(segs, mhi, mlo, mk, t3, N, h, w, 4 * cout, passes, bias=mb, act=ACT_RELU, out=out,
                                      out_place=(2 * h, 2 * w, 2, 0, 2, 0), res_post=post, out_planes=op, tag='img_tc',
                                      acc_scale=msc, planes_fmt=fmt, phase_cout=cout,
                                      flops=2.0 * N * h * w * cout * cin_t * 25)
                    x, xp = out, op
                    h, w = 2 * h, 2 * w
                    continue
                for py in range(2):
                    for px in range(2):
                        ops.conv_tc_dense(segs, hi, lo, k, ops.taps_convT_phase(py, px), N, h, w, cout, passes, bias=bd,
                                          act=ACT_RELU, out=out, out_place=(2 * h, 2 * w, 2, py, 2, px), res_post=post,
                                          out_planes=op, tag='img_tc', acc_scale=sc, planes_fmt=fmt)
                x, xp = out, op
            h, w = 2 * h, 2 * w
        wp, bp = P['pred']
        if not concat and P['pred_pw'] is not None and x.shape[-1] == P['pred_pw'].shape[1]:
            return ops.pw_conv_fwd(Seg(x), P['pred_pw'], bp, N, h, w, 1, act=ACT_SIGMOID)                  # unet.py:179
        segs = [Seg(x)] + ([Seg(head)] if concat else [])
        img, _, _, _ = ops.conv(segs, wp, bp, N, h, w, h, w, 1, ops.taps_conv(1, 0), act=ACT_SIGMOID)      # unet.py:179
        return img

    # ----------------------------------------------------------------- image decoder (fp32 kernels)
    def _image_decoder(self, P, head, blocks, N, h, w):
        ne = self.num_encoders
        concat = self.skip_type != 'sum'
        base = self.base_num_channels
        cmax = base * 2 ** ne
        blocks = [b if b.is_contiguous() else b.contiguous() for b in blocks]
        x = blocks[-1]
        t3 = ops.taps_conv(3, 1)
        nres = self.num_residual_blocks
        for j in range(nres):                                                      # submodules.py:157-172
            w1, b1, w2, b2 = P['res%d' % j]
            post = blocks[ne - 1] if (j == nres - 1 and not concat) else None      # skip_sum, unet.py:175
            if self.norm == 'IN':
                # conv1 -> InstanceNorm (per-sample statistics) -> ReLU -> conv2 -> InstanceNorm -> + x -> ReLU: the
                # statistics come out of the conv epilogue, IN + ReLU are applied by the next conv's loader
                t, _, st1, _ = ops.conv([Seg(x)], w1, b1, N, h, w, h, w, cmax, t3, want_stats=True)
                m1, r1 = ops.in_finalize(st1, h * w)
                y2, _, st2, _ = ops.conv([Seg(t, mean=m1, rstd=r1, relu=True)], w2, b2, N, h, w, h, w, cmax, t3, want_stats=True)
                m2, r2 = ops.in_finalize(st2, h * w)
                x = ops.norm_act_add(ops.norm_act_add(y2, m2, r2, relu=False, res=x), relu=True, res=post)
                continue
            t, _, _, _ = ops.conv([Seg(x)], w1, b1, N, h, w, h, w, cmax, t3, act=ACT_RELU)
            x, _, _, _ = ops.conv([Seg(t)], w2, b2, N, h, w, h, w, cmax, t3, act=ACT_RELU, res_pre=x, res_post=post)
        for i in range(ne):                                                        # unet.py:175-176
            wd, bd, cout = P['dec%d' % i]
            skip_next = blocks[ne - i - 2] if i < ne - 1 else head
            post = None if concat else skip_next
            segs = [Seg(x)] + ([Seg(blocks[ne - i - 1])] if concat else [])
            if self.use_upsample_conv:                                             # submodules.py:83-93
                if concat:
                    raise NotImplementedError('skip_type=concat with UpsampleConvLayer')
                xu = ops.bilinear_up2(x)
                x, _, _, _ = ops.conv([Seg(xu)], wd, bd, N, 2 * h, 2 * w, 2 * h, 2 * w, cout, ops.taps_conv(5, 2),
                                      act=ACT_RELU, res_post=post)
            else:                                                                  # submodules.py:53-62
                out = torch.empty((N, 2 * h, 2 * w, cout), device=x.device, dtype=torch.float32)
                for py in range(2):
                    for px in range(2):
                        ops.conv(segs, wd, bd, N, h, w, h, w, cout, ops.taps_convT_phase(py, px), act=ACT_RELU,
                                 out=out, out_place=(2 * h, 2 * w, 2, py, 2, px), res_post=post)
                x = out
            h, w = 2 * h, 2 * w
        wp, bp = P['pred']
        segs = [Seg(x)] + ([Seg(head)] if concat else [])
        img, _, _, _ = ops.conv(segs, wp, bp, N, h, w, h, w, 1, ops.taps_conv(1, 0), act=ACT_SIGMOID)  # unet.py:179
        return img

    # -------------------------------------------------------------------------- tcgen05 launches
    def _enc_conv_tc(self, e, in_planes, N, h_in, w_in, cout, out_hi, out_lo, passes, rows_out=None, tall=0):
        """conv5x5 stride 2 pad 2 + folded BN + ReLU (submodules.py:107,111) through parity views.
        rows_out: rows per image of the output planes (row-stacked level; default dense).  tall = rows per image of the
        INPUT buffer (= 2 * rows_out): run on the tall view of both (one image of N * rows rows, masked zero rows)."""
        tcw = e['tc']
        hi, lo = in_planes
        d = ConvTc()
        oh, ow = h_in // 2, w_in // 2
        rows_out = rows_out or oh
        if tall:
            # in_planes are [:, :h_in] slices of [N, tall, w_in, C] buffers whose extra rows are zero
            assert tall == 2 * rows_out and hi.stride(0) == tall * w_in * hi.shape[-1]
            hi = hi.as_strided((1, N * tall, w_in, hi.shape[-1]), (N * hi.stride(0), hi.stride(1), hi.stride(2), 1))
            lo = lo.as_strided((1, N * tall, w_in, lo.shape[-1]), (N * lo.stride(0), lo.stride(1), lo.stride(2), 1))
        taps = []
        if tcw['fold']:
            for py in range(2):
                ops.parity_view(d.views[py], hi, lo, py, 0, fold_x=True)
            d.n_views = 2
            for ky in range(5):
                for j in range(3):
                    taps.append(((ky - 2) // 2, j - 1, ky % 2, ky * 3 + j))
            cin_eff = 64
        else:
            for py in range(2):
                for px in range(2):
                    ops.parity_view(d.views[py * 2 + px], hi, lo, py, px)
            d.n_views = 4
            for ky in range(5):
                for kx in range(5):
                    taps.append(((ky - 2) // 2, (kx - 2) // 2, (ky % 2) * 2 + (kx % 2), ky * 5 + kx))
            cin_eff = hi.shape[-1]
        d.nseg = 1
        d.seg_C[0], d.seg_view0[0], d.seg_koff[0] = cin_eff, 0, 0
        d.k_per_tap, d.n_w_taps, d.w_rows = tcw['k_per_tap'], tcw['T'], tcw['hi'].shape[0]
        d.w_hi, d.w_lo, d.bias = ops._p(tcw['hi']), ops._p(tcw['lo']), ops._p(e['bias'])
        if tall:
            d.N, d.OH, d.OW, d.Cout = 1, N * rows_out, ow, cout
            d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = N * rows_out, ow, 1, 0, 1, 0
            d.row_period, d.rows_valid = rows_out, oh
        else:
            d.N, d.OH, d.OW, d.Cout = N, oh, ow, cout
            d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = rows_out, ow, 1, 0, 1, 0
        d.out_hi, d.out_lo, d.ld_planes = ops._p(out_hi), ops._p(out_lo), cout
        # `passes` = the mode's pass count (decides the OUTPUT plane format); a layer may consume another format (tcw['passes'])
        d.epilogue, d.act, d.passes, d.bw_log2 = EPI_LINEAR, ACT_RELU, tcw.get('passes', passes), ops.pick_bw_log2(ow, d.OH)
        d.acc_scale, d.planes_fmt = tcw.get('sc', 0.0), (ops.PLANES_HF8 if passes == 2 else ops.PLANES_BF16)
        d.ntaps = len(taps)
        for t, (dy, dx, v, wi) in enumerate(taps):
            d.dy[t], d.dx[t], d.view[t], d.widx[t] = dy, dx, v, wi
        ops.conv_tc(d, tag='enc_tc', device=out_hi.device)

    def _gru_tc(self, e, x_planes, h_prev, h_planes, N, oh, ow, C, passes):
        """ConvGRU cell (submodules.py:255-273) as two tcgen05 launches: [update, reset] gates (epilogue writes
        update and the bf16 planes of prev_state*reset), then the out gate over [x, prev_state*reset] whose
        epilogue blends h' = h*(1-u) + tanh(.)*u."""
        tcw = e['gru_tc']
        dev = x_planes[0].device
        taps = ops.taps_conv(3, 1)

        def base(d, second):
            ops.dense_view(d.views[0], x_planes[0], x_planes[1])
            d.n_views, d.nseg = 1, 1
            d.seg_C[0], d.seg_view0[0], d.seg_koff[0] = C, 0, 0
            if second is not None:
                ops.dense_view(d.views[1], second[0], second[1])
                d.n_views, d.nseg = 2, 2
                d.seg_C[1], d.seg_view0[1], d.seg_koff[1] = C, 1, C
            d.k_per_tap, d.n_w_taps = tcw['k_per_tap'], 9
            d.N, d.OH, d.OW = N, oh, ow
            d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = oh, ow, 1, 0, 1, 0
            d.ldo, d.ld_planes = C, C
            d.act, d.passes, d.bw_log2 = ACT_NONE, passes, ops.pick_bw_log2(ow, oh)
            d.planes_fmt = ops.PLANES_HF8 if passes == 2 else ops.PLANES_BF16
            d.ntaps = len(taps)
            for t, (dy, dx, wi) in enumerate(taps):
                d.dy[t], d.dx[t], d.view[t], d.widx[t] = dy, dx, 0, wi

        upd = torch.empty((N, oh, ow, C), device=dev, dtype=torch.float32)
        hr = (torch.empty((N, oh, ow, C), device=dev, dtype=torch.bfloat16),
              torch.empty((N, oh, ow, C), device=dev, dtype=torch.bfloat16))
        d = ConvTc()
        base(d, h_planes)
        d.w_hi, d.w_lo, d.w_rows, d.bias = ops._p(tcw['ur_hi']), ops._p(tcw['ur_lo']), tcw['ur_hi'].shape[0], ops._p(e['gru_ur_b'])
        d.aux0, d.out, d.out_hi, d.out_lo = ops._p(h_prev), ops._p(upd), ops._p(hr[0]), ops._p(hr[1])
        d.Cout, d.epilogue, d.acc_scale = 2 * C, EPI_GRU_UR, tcw.get('ur_sc', 0.0)
        ops.conv_tc(d, tag='gru_tc', device=dev)
        h = torch.empty((N, oh, ow, C), device=dev, dtype=torch.float32)
        hh = torch.empty((N, oh, ow, C), device=dev, dtype=torch.bfloat16)
        hl = torch.empty_like(hh)
        d = ConvTc()
        base(d, hr if h_planes is not None else None)          # prev_state = 0  =>  prev_state*reset = 0: skip that K half
        d.w_hi, d.w_lo, d.w_rows, d.bias = ops._p(tcw['o_hi']), ops._p(tcw['o_lo']), tcw['o_hi'].shape[0], ops._p(e['gru_o_b'])
        d.aux0, d.aux1, d.out, d.out_hi, d.out_lo = ops._p(h_prev), ops._p(upd), ops._p(h), ops._p(hh), ops._p(hl)
        d.Cout, d.epilogue, d.acc_scale = C, EPI_GRU_OUT, tcw.get('o_sc', 0.0)
        ops.conv_tc(d, tag='gru_tc', device=dev)
        return h, hh, hl

    def _lstm_tc(self, e, x_planes, h_planes, c_prev, N, oh, ow, C, passes, rows=None, out_planes=None):
        """Fused ConvLSTM cell (submodules.py:190-230): gates GEMM + sigma/tanh + state update.
        rows > oh: row-stacked level -- every tensor is a full [N, rows, ow, C] buffer (zero rows in the planes) and the
        launch sees ONE image of N * rows rows; h / c come back as full buffers too (rows >= oh of an image unwritten)."""
        tcw = e['lstm_tc']
        d = ConvTc()
        n_img, img_rows = N, oh
        if rows is not None and rows > oh:
            tall = lambda t: t.view(1, N * rows, ow, t.shape[-1])
            x_planes = (tall(x_planes[0]), tall(x_planes[1]))
            if h_planes is not None:
                h_planes = (tall(h_planes[0]), tall(h_planes[1]))
            d.row_period, d.rows_valid = rows, oh
            N, oh = 1, N * rows
        ops.dense_view(d.views[0], x_planes[0], x_planes[1])
        d.n_views, d.nseg = 1, 1
        d.seg_C[0], d.seg_view0[0], d.seg_koff[0] = C, 0, 0
        if h_planes is not None:
            ops.dense_view(d.views[1], h_planes[0], h_planes[1])
            d.n_views, d.nseg = 2, 2
            d.seg_C[1], d.seg_view0[1], d.seg_koff[1] = C, 1, C
        d.k_per_tap, d.n_w_taps, d.w_rows = tcw['k_per_tap'], 9, tcw['hi'].shape[0]
        d.w_hi, d.w_lo, d.bias = ops._p(tcw['hi']), ops._p(tcw['lo']), ops._p(e['lstm_b'])
        dev = x_planes[0].device
        full = (n_img, rows, ow, C) if d.row_period else (N, oh, ow, C)
        h = torch.empty(full, device=dev, dtype=torch.float32)
        c = torch.empty_like(h)
        if out_planes is not None:
            hh, hl = out_planes
        else:
            hh = torch.empty(full, device=dev, dtype=torch.bfloat16)
            hl = torch.empty_like(hh)
        d.aux0, d.out, d.out2 = ops._p(c_prev), ops._p(h), ops._p(c)
        d.out_hi, d.out_lo, d.ld_planes = ops._p(hh), ops._p(hl), C
        d.N, d.OH, d.OW, d.Cout = N, oh, ow, 4 * C
        d.OHf, d.OWf, d.osy, d.ooy, d.osx, d.oox = oh, ow, 1, 0, 1, 0
        d.ldo = C
        d.epilogue, d.act, d.passes, d.bw_log2 = EPI_LSTM, ACT_NONE, passes, ops.pick_bw_log2(ow, oh)
        d.acc_scale, d.planes_fmt = tcw.get('sc', 0.0), (ops.PLANES_HF8 if passes == 2 else ops.PLANES_BF16)
        taps = ops.taps_conv(3, 1)
        d.ntaps = len(taps)
        for t, (dy, dx, wi) in enumerate(taps):
            d.dy[t], d.dx[t], d.view[t], d.widx[t] = dy, dx, 0, wi
        ops.conv_tc(d, tag='lstm_tc', device=dev)
        return h, c, hh, hl
