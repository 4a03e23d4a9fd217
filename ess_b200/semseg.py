"""Drop-in `SemSegE2VID` segmentation decoder (reference: models/style_networks.py:9-107) with a
hand-scheduled forward AND backward on libess_b200.so kernels.

Same constructor arguments, `forward(input_dict{1,2,4,8}) -> dict{8,4,2,1}`, parameter names
(`decoder_scale_1.{0-4}.model.{0,3}.*`, `decoder_scale_1.5.model.0.*`, ... `decoder_scale_5.0.*`) and
initialisation (N(0, 0.02) conv weights inside the IN blocks, style_networks.py:152-155) as the
reference, so `utils/saver.py` checkpoints round-trip.

Execution model: the decoder is a static list of nodes (a tiny graph executor instead of autograd
tracing): `conv` nodes run the implicit-GEMM kernel whose loader applies InstanceNorm + ReLU +
nearest x2 upsampling + channel concat on the fly and whose epilogue emits per-(n,c) sum / sumsq
partials for the NEXT InstanceNorm; `mat` nodes materialise an IN(+ReLU)(+residual) result where the
graph needs it as a tensor (INSResBlock outputs, the returned out[4] / out[2]).  The backward walks the
same list in reverse: dgrad = the same gather kernel with mirrored taps and transposed weights
(per input segment, skipped for segments that need no gradient), wgrad = split-K implicit GEMM,
InstanceNorm/ReLU/upsample backward = two memory-bound passes.  One torch.autograd.Function wraps
the whole decoder and honours `requires_grad` on every parameter and input individually
(training/ess_trainer.py:133-137 freezes the parameters but still needs input gradients).
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE, EPI_LINEAR
from .ops import Seg


# --------------------------------------------------------------------------------- parameter holders
class _ReLUINSConv2d(nn.Module):      # style_networks.py:158-169
    def __init__(self, n_in, n_out):
        super().__init__()
        self.model = nn.Sequential(nn.Conv2d(n_in, n_out, 3, 1, 1, bias=True), nn.Identity(), nn.Identity())
        self.model[0].weight.data.normal_(0.0, 0.02)


class _INSResBlock(nn.Module):        # style_networks.py:172-193
    def __init__(self, c):
        super().__init__()
        self.model = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1), nn.Identity(), nn.Identity(), nn.Conv2d(c, c, 3, 1, 1),
                                   nn.Identity())
        self.model[0].weight.data.normal_(0.0, 0.02)
        self.model[3].weight.data.normal_(0.0, 0.02)


class _Conv(object):
    """`cout` = channels of the node's output tensor.  `pad` = None, or (Cout, Cin) of the PARAMETER when the executor
    runs the layer zero-padded to (cout, sum of source channels): input_index_map=True makes the first stage 258
    channels wide, which is carried as 260 (vector-width multiple) with zero weights / zero activations in the pad."""
    __slots__ = ('out', 'srcs', 'w', 'b', 'cout', 'k', 'stats', 'pad')

    def __init__(self, out, srcs, w, b, cout, k, stats, pad=None):
        self.out, self.srcs, self.w, self.b, self.cout, self.k, self.stats = out, srcs, w, b, cout, k, stats
        self.pad = pad


class _Mat(object):
    __slots__ = ('out', 'src', 'relu', 'res')

    def __init__(self, out, src, relu, res):
        self.out, self.src, self.relu, self.res = out, src, relu, res


class SemSegE2VID(nn.Module):
    def __init__(self, input_c, output_c, skip_connect=False, skip_type='sum', input_index_map=False):
        super().__init__()
        self.skip_connect = skip_connect
        self.skip_type = skip_type
        self.input_index_map = input_index_map
        self.index_coords = None
        self.input_c, self.output_c = input_c, output_c
        tch = input_c
        if skip_connect:
            if skip_type != 'concat':
                # style_networks.py:77-79: skip_sum keeps tch//2 channels but decoder_scale_2 expects tch
                raise NotImplementedError("SemSegE2VID(skip_connect=True) is only shape-consistent with "
                                          "skip_type='concat' (as in the reference)")
            self.decoder_scale_1 = nn.Sequential(*([_INSResBlock(tch) for _ in range(5)] +
                                                   [_ReLUINSConv2d(tch, tch // 2)]))
            self.decoder_scale_2 = nn.Sequential(_ReLUINSConv2d(tch, tch // 2), _ReLUINSConv2d(tch // 2, tch // 4))
            tch //= 2
            self.decoder_scale_3 = nn.Sequential(_ReLUINSConv2d(tch, tch // 2), _ReLUINSConv2d(tch // 2, tch // 2))
            tch //= 2
            self.decoder_scale_4 = nn.Sequential(_ReLUINSConv2d(tch, tch // 2))
            tch //= 2
        else:
            if input_index_map:                                  # style_networks.py:36-50: two coordinate channels
                tch += 2
            self.decoder_scale_1 = nn.Sequential(*[_INSResBlock(tch) for _ in range(3)])
            if input_index_map:
                self.decoder_scale_2 = nn.Sequential(nn.Identity(), _ReLUINSConv2d(tch, (tch - 2) // 2))
                tch = (tch - 2) // 2
            else:
                self.decoder_scale_2 = nn.Sequential(nn.Identity(), _ReLUINSConv2d(tch, tch // 2))
                tch //= 2
            self.decoder_scale_3 = nn.Sequential(nn.Identity(), _ReLUINSConv2d(tch, tch // 2))
            tch //= 2
            self.decoder_scale_4 = nn.Sequential(nn.Identity(), _ReLUINSConv2d(tch, tch // 2))
            tch //= 2
        self.decoder_scale_5 = nn.Sequential(nn.Conv2d(tch, output_c, 1, 1, 0))
        # precision mode of the 3x3 convolutions (forward + dgrad): 'fp32' (CUDA cores, exact),
        # 'bf16x3' (tcgen05, fp32-parity split) or 'bf16'; constructor signature stays the reference's
        from .e2vid import default_mode
        self.mode = default_mode()
        self._build_program()

    # ------------------------------------------------------------------------------- static graph
    def _build_program(self):
        """Tensor ids: 0 = input[8], 1 = input[4], 2 = input[2]; the rest are produced by nodes."""
        nodes, outs = [], []
        nid = [3]

        def new():
            nid[0] += 1
            return nid[0] - 1

        def conv(srcs, prefix, cout, k=3, stats=True, pad=None):
            o = new()
            nodes.append(_Conv(o, srcs, prefix + '.weight', prefix + '.bias', cout, k, stats, pad))
            return o

        def mat(src, relu, res=None):
            o = new()
            nodes.append(_Mat(o, src, relu, res))
            return o

        c = self.input_c
        x = 0
        nres = 5 if self.skip_connect else 3
        cw, pad1 = c, None                                               # width of the first stage as executed
        if self.input_index_map and not self.skip_connect:
            cw = (c + 2 + 3) // 4 * 4                                    # 258 -> 260
            pad1 = (c + 2, c + 2)
        for b in range(nres):                                            # INSResBlock
            y1 = conv([(x, 'raw', 0)], 'decoder_scale_1.%d.model.0' % b, cw, pad=pad1)
            y2 = conv([(y1, 'nr', 0)], 'decoder_scale_1.%d.model.3' % b, cw, pad=pad1)
            x = mat(y2, False, x)
        if self.skip_connect:
            y = conv([(x, 'raw', 0)], 'decoder_scale_1.5.model.0', c // 2)
            y = conv([(y, 'nr', 1), (1, 'raw', 0)], 'decoder_scale_2.0.model.0', c // 2)
            y = conv([(y, 'nr', 0)], 'decoder_scale_2.1.model.0', c // 4)
            o4 = mat(y, True)
            outs.append(o4)
            y = conv([(o4, 'raw', 1), (2, 'raw', 0)], 'decoder_scale_3.0.model.0', c // 4)
            y = conv([(y, 'nr', 0)], 'decoder_scale_3.1.model.0', c // 4)
            o2 = mat(y, True)
            outs.append(o2)
            y = conv([(o2, 'raw', 1)], 'decoder_scale_4.0.model.0', c // 8)
            o1 = conv([(y, 'nr', 0)], 'decoder_scale_5.0', self.output_c, k=1, stats=False)
            outs.append(o1)
        else:
            y = conv([(x, 'raw', 1)], 'decoder_scale_2.1.model.0', c // 2, pad=(c // 2, c + 2) if pad1 else None)
            o4 = mat(y, True)
            outs.append(o4)
            y = conv([(o4, 'raw', 1)], 'decoder_scale_3.1.model.0', c // 4)
            o2 = mat(y, True)
            outs.append(o2)
            y = conv([(o2, 'raw', 1)], 'decoder_scale_4.1.model.0', c // 8)
            o1 = conv([(y, 'nr', 0)], 'decoder_scale_5.0', self.output_c, k=1, stats=False)
            outs.append(o1)
        self._nodes, self._outs = nodes, outs

    def update_skip_dict(self, skips, x, sz_in):
        rem, scale = sz_in % x.shape[3], sz_in // x.shape[3]
        assert rem == 0
        skips[scale] = x

    def forward(self, input_dict):
        sz_in = input_dict[1].shape[3]        # input_dict[1] is used for its width only (:70)
        x8 = input_dict[8]
        ops.require_cuda_any(x8)
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        if self.skip_connect:
            ins = (x8, input_dict[4], input_dict[2])
        else:
            ins = (x8, None, None)
        if self.input_index_map and not self.skip_connect:
            # style_networks.py:90-97: x-coordinate = row index, y-coordinate = column index, cached on the module
            if self.index_coords is None or self.index_coords.shape[0] != x8.shape[0] or \
                    self.index_coords.shape[2:] != x8.shape[2:] or self.index_coords.device != x8.device:
                xc = torch.arange(x8.size(2), device=x8.device, dtype=torch.float)
                yc = torch.arange(x8.size(3), device=x8.device, dtype=torch.float)
                self.index_coords = torch.stack(torch.meshgrid([xc, yc], indexing='ij'), 0)[None].repeat(x8.size(0), 1, 1, 1)
        with ops.on_device_of(x8):
            res = _DecoderFn.apply(self, names, ins[0], ins[1], ins[2], *params)
        out = {8: x8}
        for t in res:
            self.update_skip_dict(out, t, sz_in)
        return out


def _nhwc(t):
    if t is None:
        return None
    v = t.detach().float().permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def _taps(k):
    return ops.taps_conv(k, k // 2)


def _weight(nd, w):
    """The node's weight as executed: the parameter itself, or zero-padded to (nd.cout, padded Cin) -- see _Conv.pad"""
    w = w.detach()
    if nd.pad is None:
        return w
    cout, cin = nd.pad
    cin_p = (cin + 3) // 4 * 4
    wp = torch.zeros((nd.cout, cin_p) + tuple(w.shape[2:]), device=w.device, dtype=torch.float32)
    wp[:cout, :cin] = w
    return wp


def _bias(nd, b):
    b = b.detach().float()
    if nd.pad is None or b.numel() == nd.cout:
        return b.contiguous()
    bp = torch.zeros((nd.cout,), device=b.device, dtype=torch.float32)
    bp[:b.numel()] = b
    return bp


def _packed32(module, w, nd, c_off=None, cs=None):
    """ops.pack_weight of the node's (padded) weight for the fp32 CUDA-core kernels -- forward layout, or, with
    (c_off, cs), the swap_io layout of the input-channel slice used by the input gradient -- cached per parameter
    version like the tensor-core packs (the UDA iteration runs the decoder 5x forward / 3x backward per step)."""
    cache = module.__dict__.setdefault('_wcache', {})
    key = (w.data_ptr(), 'fp32', c_off, cs)
    ent = cache.get(key)
    if ent is None or ent[0] != w._version or ent[1] != tuple(w.shape):
        wx = _weight(nd, w)
        packed = ops.pack_weight(wx) if c_off is None else ops.pack_weight(wx[:, c_off:c_off + cs].contiguous(), swap_io=True)
        ent = (w._version, tuple(w.shape), packed)
        cache[key] = ent
    return ent[2]


def _packed_tc(module, w, **kw):
    """pack_weight_tc(w, **kw) cached on the module per (parameter storage, version, options): the UDA iteration
    runs the decoder five times forward and three times backward between two optimizer steps."""
    cache = module.__dict__.setdefault('_wcache', {})
    key = (w.data_ptr(), tuple(sorted(kw.items())))
    ent = cache.get(key)
    if ent is None or ent[0] != w._version or ent[1] != tuple(w.shape):
        ent = (w._version, tuple(w.shape), ops.pack_weight_tc(w, **kw))
        cache[key] = ent
    return ent[2]


def _bias_grad(nd, gy, dev):
    """Bias gradient of a conv node.  A bias that feeds an InstanceNorm (every conv of the IN blocks,
    style_networks.py:162-163,174-183) is cancelled by the mean subtraction: the incoming gradient gy is the
    output of the InstanceNorm backward, whose per-(n, c) sum over pixels is identically zero (the reference
    produces ~1e-10 rounding noise there).  Those gradients are returned as exact zeros instead of being
    summed over all pixels; only a conv whose output is not normalised has a real bias gradient."""
    if nd.stats:
        return torch.zeros((nd.cout,), device=dev, dtype=torch.float32)
    return ops.colsum(gy, nd.cout)


def _pw_ok(nd, segs, cin_total):
    """1x1 conv with one non-upsampled 32/64-channel source and <= 16 outputs -> pw_conv.cu kernels."""
    return nd.k == 1 and not nd.stats and len(segs) == 1 and segs[0].ups == 0 and cin_total in (32, 64) and \
        nd.cout <= 16 and segs[0].t.shape[-1] == cin_total


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, names, x8, x4, x2, *params):
        nodes = module._nodes
        P = dict(zip(names, params))
        T = {0: _nhwc(x8), 1: _nhwc(x4), 2: _nhwc(x2)}
        if module.input_index_map and not module.skip_connect:
            # x = cat([x, index_coords]) (style_networks.py:97), carried with zero channels up to a multiple of 4
            cw = nodes[0].cout
            x0 = torch.zeros(tuple(T[0].shape[:3]) + (cw,), device=x8.device, dtype=torch.float32)
            x0[..., :module.input_c] = T[0]
            x0[..., module.input_c:module.input_c + 2] = module.index_coords.permute(0, 2, 3, 1)
            T[0] = x0
        S = {}
        tc = module.mode != 'fp32'
        passes = 1 if module.mode == 'bf16' else 3      # 'f16f8' (an encoder mode) = bf16x3 here
        ctx.mode = module.mode
        # operand planes of every tensor-core conv are kept for its weight gradient (same bytes as the fp32
        # activation; ~1.7 GB at B=8 DSEC) instead of being re-created in the backward pass
        keep_planes = any(ctx.needs_input_grad[5:])
        ctx.planes = {}
        probe = module.__dict__.get('_probe')    # test hook (teacher forcing, tests/test_gpu_teacher.py); None in use
        for nd in nodes:
            if isinstance(nd, _Conv):
                first = T[nd.srcs[0][0]]
                N = first.shape[0]
                ups0 = nd.srcs[0][2]
                H, W = first.shape[1] << ups0, first.shape[2] << ups0
                segs = []
                for (sid, xf, ups) in nd.srcs:
                    if xf == 'nr':
                        segs.append(Seg(T[sid], ups=ups, mean=S[sid][0], rstd=S[sid][1], relu=True))
                    else:
                        segs.append(Seg(T[sid], ups=ups))
                w = P[nd.w]
                bias = _bias(nd, P[nd.b])
                cin_total = w.shape[1] if nd.pad is None else (nd.pad[1] + 3) // 4 * 4
                if tc and nd.k == 3 and nd.pad is None and cin_total % 64 == 0 and nd.cout in (32, 64, 128, 256):
                    # tensor-core path: operand planes = the transformed (IN/ReLU/upsample/concat) input
                    hi = torch.empty((N, H, W, cin_total), device=first.device, dtype=torch.bfloat16)
                    lo = torch.empty_like(hi)
                    c_off = 0
                    for sg in segs:
                        ops.split_bf16(sg, N, H, W, hi, lo, c_off)
                        c_off += sg.C if sg.C is not None else sg.t.shape[-1]
                    w_hi, w_lo, kinp = _packed_tc(module, w)
                    y = ops.conv_tc_dense((hi, lo), w_hi, w_lo, kinp, _taps(nd.k), N, H, W, nd.cout, passes, bias=bias,
                                          tag='seg_fwd')
                    if keep_planes:
                        ctx.planes[nd.out] = (hi, lo)
                    del hi, lo
                    if probe is not None:
                        y = probe.fwd(nd.out, y)
                    T[nd.out] = y
                    if nd.stats:
                        S[nd.out] = ops.in_stats(y)
                elif _pw_ok(nd, segs, cin_total):
                    # 1x1 classifier (style_networks.py:34,88): HBM-bound, dedicated streaming kernel
                    T[nd.out] = ops.pw_conv_fwd(segs[0], w.detach().float().reshape(nd.cout, cin_total), bias, N, H, W,
                                                nd.cout)
                    if probe is not None:
                        T[nd.out] = probe.fwd(nd.out, T[nd.out])
                else:
                    wp = _packed32(module, w, nd)
                    y, _, st, _ = ops.conv(segs, wp, bias, N, H, W, H, W, nd.cout, _taps(nd.k), epilogue=EPI_LINEAR,
                                           act=ACT_NONE, want_stats=nd.stats)
                    if probe is not None:
                        y2 = probe.fwd(nd.out, y)
                        if y2 is not y:
                            y, st = y2, None
                    T[nd.out] = y
                    if nd.stats:
                        S[nd.out] = ops.in_finalize(st, H * W) if st is not None else ops.in_stats(y)
            else:
                src = T[nd.src]
                T[nd.out] = ops.norm_act_add(src, S[nd.src][0], S[nd.src][1], relu=nd.relu,
                                             res=T[nd.res] if nd.res is not None else None)
                if probe is not None:
                    T[nd.out] = probe.fwd(nd.out, T[nd.out])
        ctx.module, ctx.names = module, names
        ctx.T, ctx.S = T, S
        ctx.set_materialize_grads(False)     # unused outputs (out[4], out[2] in the supervised step) stay None
        ctx.save_for_backward(*params)
        outs = tuple(ops.as_nchw(T[o]) for o in module._outs)
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        with ops.on_device_of(ctx.T[0]):
            return _DecoderFn._backward(ctx, *gouts)

    @staticmethod
    def _backward(ctx, *gouts):
        module, names = ctx.module, ctx.names
        nodes = module._nodes
        params = ctx.saved_tensors
        P = dict(zip(names, params))
        T, S = ctx.T, ctx.S
        nig = ctx.needs_input_grad          # (module, names, x8, x4, x2, *params)
        need_in = {0: nig[2], 1: nig[3], 2: nig[4]}
        need_p = {n: nig[5 + i] for i, n in enumerate(names)}

        # which tensors carry gradient
        need = dict(need_in)
        for nd in nodes:
            if isinstance(nd, _Conv):
                need[nd.out] = need_p[nd.w] or need_p[nd.b] or any(need[s[0]] for s in nd.srcs)
            else:
                need[nd.out] = need[nd.src] or (nd.res is not None and need[nd.res])

        G = {}   # tensor id -> [grad NHWC fp32 tensor | None, owned, bf16 (hi, lo) planes | None]
        probe = module.__dict__.get('_probe')
        sink = module.__dict__.get('_grad_sink')   # ess_b200.dp.GradBucket(module=...): overlapped gradient exchange
        conv_of = {nd.out: nd for nd in nodes if isinstance(nd, _Conv)}

        def planes_spec(tid):
            """How the InstanceNorm backward should emit the gradient of conv output `tid`: (pitch of the bf16
            hi/lo planes or 0, fp32 copy still needed).  When the producing conv runs its dgrad and wgrad on the
            tcgen05 kernels the gradient is written ONLY in their operand format (no fp32 round trip)."""
            nd = conv_of.get(tid)
            if nd is None or ctx.mode == 'fp32' or nd.k != 3 or nd.cout % 32:
                return 0, True
            ok = True
            if need_p[nd.w] or need_p[nd.b]:
                ok = P[nd.w].shape[1] % 64 == 0 and nd.cout <= 256 and (nd.stats or not need_p[nd.b])
            for (sid, _, _) in nd.srcs:
                if need[sid] and T[sid].shape[-1] not in (64, 128, 256):
                    ok = False
            return (nd.cout + 63) // 64 * 64, not ok

        def add_grad(tid, g, owned, planes=None):
            if not need.get(tid, False):
                return
            if tid not in G:
                G[tid] = [g, owned, planes]
            else:
                cur = G[tid]
                if cur[0] is None or g is None:
                    raise RuntimeError('internal: plane-only gradient of a tensor with two consumers')
                if not cur[1]:
                    cur[0] = cur[0].clone()
                    cur[1] = True
                cur[2] = None
                ops.norm_act_add(cur[0], res=g, out=cur[0])

        def in_bwd(tid, dA, y, mean, rstd, relu, ups=0):
            ld, want32 = planes_spec(tid)
            if not ld:
                return ops.in_backward(dA, y, mean, rstd, relu=relu, ups=ups), None
            return ops.in_backward(dA, y, mean, rstd, relu=relu, ups=ups, planes_ld=ld, want_fp32=want32)

        for o, g in zip(module._outs, gouts):
            if g is not None:
                add_grad(o, _nhwc(g), False)

        grads = {}
        for nd in reversed(nodes):
            if nd.out not in G:
                continue
            gy, _, gy_planes = G.pop(nd.out)
            if probe is not None:
                gy, gy_planes = probe.bwd(nd.out, gy, gy_planes)
            if isinstance(nd, _Mat):
                y = T[nd.src]
                mean, rstd = S[nd.src]
                if need[nd.src]:
                    g32, gpl = in_bwd(nd.src, gy, y, mean, rstd, nd.relu)
                    add_grad(nd.src, g32, True, gpl)
                if nd.res is not None:
                    add_grad(nd.res, gy, False)   # ownership not transferred: gy may alias a caller tensor
                continue
            first = T[nd.srcs[0][0]]
            N = first.shape[0]
            ups0 = nd.srcs[0][2]
            H, W = first.shape[1] << ups0, first.shape[2] << ups0
            w = P[nd.w]
            taps = _taps(nd.k)
            tc = ctx.mode != 'fp32' and nd.k == 3 and nd.cout % 32 == 0 and nd.pad is None
            passes = 1 if ctx.mode == 'bf16' else 3
            gplanes = gy_planes

            def dy_planes():
                """bf16 hi/lo planes of dY (channel-padded to a multiple of 64 with zeros), built once per node"""
                kinp = (nd.cout + 63) // 64 * 64
                pl = (torch.empty((N, H, W, kinp), device=gy.device, dtype=torch.bfloat16),
                      torch.empty((N, H, W, kinp), device=gy.device, dtype=torch.bfloat16))
                ops.split_bf16(Seg(gy), N, H, W, pl[0], pl[1], 0, c_pad=kinp)
                return pl

            if need_p[nd.w] or need_p[nd.b]:
                segs = []
                for (sid, xf, ups) in nd.srcs:
                    if xf == 'nr':
                        segs.append(Seg(T[sid], ups=ups, mean=S[sid][0], rstd=S[sid][1], relu=True))
                    else:
                        segs.append(Seg(T[sid], ups=ups))
                cin_total = w.shape[1] if nd.pad is None else (nd.pad[1] + 3) // 4 * 4
                if _pw_ok(nd, segs, cin_total):
                    dw, db = ops.pw_conv_wgrad(segs[0], gy, want_w=need_p[nd.w], want_b=need_p[nd.b])
                    if need_p[nd.w]:
                        grads[nd.w] = dw.view(w.shape)
                    if need_p[nd.b]:
                        grads[nd.b] = db
                elif tc and cin_total % 64 == 0 and nd.cout <= 256:
                    # tensor-core wgrad: the conv's operand planes kept by the forward pass (re-created if absent)
                    kept = ctx.planes.pop(nd.out, None)
                    if kept is not None:
                        hi, lo = kept
                    else:
                        hi = torch.empty((N, H, W, cin_total), device=w.device, dtype=torch.bfloat16)
                        lo = torch.empty_like(hi)
                        co = 0
                        for sg in segs:
                            ops.split_bf16(sg, N, H, W, hi, lo, co)
                            co += sg.C if sg.C is not None else sg.t.shape[-1]
                    del kept
                    if gplanes is None:
                        gplanes = dy_planes()
                    if need_p[nd.w]:
                        grads[nd.w] = ops.wgrad_tc((hi, lo), gplanes, cin_total, nd.cout, taps, N, H, W, passes).view(w.shape)
                    del hi, lo
                    if need_p[nd.b]:
                        grads[nd.b] = _bias_grad(nd, gy, w.device)
                else:
                    dw, db = ops.wgrad(segs, gy, N, H, W, H, W, nd.cout, taps, want_bias=need_p[nd.b])
                    if nd.pad is not None:            # drop the zero-padded rows / columns
                        dw = dw[:nd.pad[0], :nd.pad[1]].contiguous()
                        db = db[:nd.pad[0]].contiguous() if db is not None else None
                    if need_p[nd.w]:
                        grads[nd.w] = dw.view(w.shape)
                    if need_p[nd.b]:
                        grads[nd.b] = db
                if sink is not None:       # hand the finished gradients over now: their exchange overlaps what follows
                    for key in (nd.w, nd.b):
                        if key in grads and sink(key, grads[key]):
                            del grads[key]
            c_off = 0
            dtaps = [(-dy, -dx, wi) for (dy, dx, wi) in taps]
            for (sid, xf, ups) in nd.srcs:
                src = T[sid]
                cs = src.shape[-1]
                if need[sid]:
                    wseg = w.detach()[:, c_off:c_off + cs].contiguous() if nd.pad is None else None
                    if nd.k == 1 and len(nd.srcs) == 1 and cs in (32, 64) and nd.cout <= 16 and not ups:
                        dA = ops.pw_conv_dgrad(gy, wseg.float().reshape(nd.cout, cs), cs)
                    elif tc and cs in (64, 128, 256):
                        kinp = (nd.cout + 63) // 64 * 64
                        if gplanes is None:
                            gplanes = dy_planes()
                        cache = module.__dict__.setdefault('_wcache', {})
                        key = (w.data_ptr(), 'dgrad', c_off, cs, kinp)
                        ent = cache.get(key)
                        if ent is None or ent[0] != w._version:
                            ent = (w._version, ops.pack_weight_tc(wseg, swap_io=True, kin_pad=kinp))
                            cache[key] = ent
                        w_hi, w_lo, _ = ent[1]
                        dA = ops.conv_tc_dense(gplanes, w_hi, w_lo, kinp, dtaps, N, H, W, cs, passes, tag='seg_dgrad')
                    else:
                        wp = _packed32(module, w, nd, c_off, cs)
                        dA, _, _, _ = ops.conv([Seg(gy)], wp, None, N, H, W, H, W, cs, dtaps)
                    if xf == 'nr':
                        g32, gpl = in_bwd(sid, dA, src, S[sid][0], S[sid][1], True, ups=ups)
                        add_grad(sid, g32, True, gpl)
                    elif ups:
                        add_grad(sid, ops.upsample2_bwd(dA, H >> 1, W >> 1, cs), True)
                    else:
                        add_grad(sid, dA, True)
                c_off += cs

        def out_grad(tid, ref):
            if tid in G and ref is not None:
                g = G[tid][0]
                if tid == 0 and g.shape[-1] != module.input_c:      # input_index_map: drop coordinate / pad channels
                    g = g[..., :module.input_c].contiguous()
                return ops.as_nchw(g).to(ref.dtype)
            return None

        gx8 = out_grad(0, None if not nig[2] else T[0])
        gx4 = out_grad(1, None if not nig[3] else T[1])
        gx2 = out_grad(2, None if not nig[4] else T[2])
        gparams = tuple(grads.get(n) for n in names)
        return (None, None, gx8, gx4, gx2) + gparams
