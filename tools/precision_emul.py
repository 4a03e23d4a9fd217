"""CPU emulation of reduced-pass tensor-core operand schemes inside the oracle (no GPU needed): which operand
splits keep the logits of the T-window unroll + decoder within the 1e-3 contract of the fp32 reference?

    python tools/precision_emul.py [T] [H] [W]

Every F.conv2d of the E2VID encoder (and optionally the decoder) is replaced by the sum of convolutions over ROUNDED
operands, accumulated in fp32 -- exactly what the tensor pipe computes (products of low-precision operands are exact in
the fp32 accumulator).  Schemes:
  bf16x3 : Ah*Wh + Al*Wh + Ah*Wl          (bf16 hi/lo; 3 bf16 passes)                      -- shipped
  f16x2w : A16*W16 + A16*Wl16             (weights split, activations single fp16; 2 passes)
  f16f8  : A16*W16 + e4m3(A)*e4m3(Wl) + e4m3(Al)*e4m3(W)   (fp16 pass + two fp8 cross terms = 2 pass-equivalents)
"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
sys.path.insert(0, __import__('os').path.join(sys.path[0], 'tests'))
from oracle import ess_oracle as O  # noqa: E402
from helpers import E2VID_CFG, make_e2vid, make_events, make_labels, make_semseg, sd_cpu  # noqa: E402

_conv = F.conv2d
STATS = {}


def e4m3(x, log2scale):
    s = 2.0 ** log2scale
    return (x * s).clamp(-448, 448).to(torch.float8_e4m3fn).float() / s


def pow2_scale(t, target=256.0):
    m = float(t.abs().max())
    import math
    return math.floor(math.log2(target / m)) if m > 0 else 0


def make_conv(scheme, a8=4):
    def conv(x, w, b=None, **kw):
        if scheme == 'fp32':
            return _conv(x, w, b, **kw)
        STATS['amax'] = max(STATS.get('amax', 0.0), float(x.abs().max()))
        if scheme == 'bf16x3':
            xh, wh = x.bfloat16().float(), w.bfloat16().float()
            xl, wl = (x - xh).bfloat16().float(), (w - wh).bfloat16().float()
            y = _conv(xh, wh, None, **kw) + _conv(xl, wh, None, **kw) + _conv(xh, wl, None, **kw)
        elif scheme == 'f16x2w':
            xh, wh = x.half().float(), w.half().float()
            wl = (w - wh).half().float()
            y = _conv(xh, wh, None, **kw) + _conv(xh, wl, None, **kw)
        elif scheme == 'f16f8':
            xh, wh = x.half().float(), w.half().float()
            xl, wl = x - xh, w - wh
            w8 = pow2_scale(w)
            y = _conv(xh, wh, None, **kw) + _conv(e4m3(x, a8), e4m3(wl, w8 + 11), None, **kw) + \
                _conv(e4m3(xl, a8 + 11), e4m3(w, w8), None, **kw)
        elif scheme == 'bf16f8':
            xh, wh = x.bfloat16().float(), w.bfloat16().float()
            xl, wl = x - xh, w - wh
            w8 = pow2_scale(w)
            y = _conv(xh, wh, None, **kw) + _conv(e4m3(x, a8), e4m3(wl, w8 + 8), None, **kw) + \
                _conv(e4m3(xl, a8 + 8), e4m3(w, w8), None, **kw)
        else:
            raise ValueError(scheme)
        return y if b is None else y + b.view(1, -1, 1, 1)
    return conv


def run(scheme, where, data, labels, e_sd, d_sd, T, C, K, a8=4):
    STATS.clear()
    F.conv2d = make_conv(scheme, a8) if where in ('encoder', 'both') else _conv
    try:
        with torch.no_grad():
            _, _, lat = O.encoder_unroll(e_sd, E2VID_CFG, data, T, C)
        F.conv2d = make_conv(scheme, a8) if where in ('decoder', 'both') else _conv
        with torch.no_grad():
            pred = O.semseg_forward(d_sd, lat)
    finally:
        F.conv2d = _conv
    return lat, pred[1]


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    B, C, K = 2, 5, 11
    m = make_e2vid(mode='fp32')
    e_sd = sd_cpu(m)
    d_sd = sd_cpu(make_semseg(K))
    data = make_events(B, T, C, H, W)
    labels = make_labels(B, H, W, K)
    lat0, log0 = run('fp32', 'both', data, labels, e_sd, d_sd, T, C, K)
    e64 = {k: v.double() for k, v in e_sd.items()}
    d64 = {k: v.double() for k, v in d_sd.items()}
    lat64, log64 = run('fp32', 'both', data.double(), labels, e64, d64, T, C, K)
    print('T=%d %dx%d B=%d;  fp32 reference vs fp64: latent8 %.2e logits %.2e' % (T, H, W, B, rel(lat0[8], lat64[8]), rel(log0, log64)))
    for scheme, where, a8 in (('bf16x3', 'both', 0), ('f16f8', 'encoder', 4), ('f16f8', 'encoder', 3), ('f16f8', 'both', 4),
                              ('bf16f8', 'encoder', 4), ('f16x2w', 'encoder', 0)):
        lat, log = run(scheme, where, data, labels, e_sd, d_sd, T, C, K, a8)
        print('%-8s in %-8s a8=%d: latent[8] %.2e  latent[2] %.2e  latent[1] %.2e  logits %.2e   (max |activation| seen %.1f)' %
              (scheme, where, a8, rel(lat[8], lat0[8]), rel(lat[2], lat0[2]), rel(lat[1], lat0[1]), rel(log, log0), STATS.get('amax', 0)))


if __name__ == '__main__':
    main()
