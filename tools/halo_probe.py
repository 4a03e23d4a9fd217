"""What bounds the narrow tcgen05 layers?  Times the head conv (5->32, N=128), encoder conv 0 (32->64, N=64) and encoder conv
1 (64->128, N=128) of one DSEC window (B=8) in place with parts of the kernel switched off (ESSB_TC_DEBUG, results garbage):
    python tools/halo_probe.py
"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ess_b200  # noqa: E402
from ess_b200 import _lib  # noqa: E402
from helpers import make_e2vid, make_events  # noqa: E402

B, T, C, H, W = 8, 3, 5, 440, 640
m = make_e2vid(mode=ess_b200.e2vid.default_mode()).cuda()
rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
data = make_events(B, T, C, H, W).cuda()


def run(label):
    for _ in range(2):
        rec.unroll(data, T, C)
    torch.cuda.synchronize()
    _lib.PROFILE, _lib.PROFILE_TAGS = [], None
    for _ in range(6):
        rec.unroll(data, T, C)
    torch.cuda.synchronize()
    acc = {}
    order = []
    for i, (tag, fl, a, b) in enumerate(_lib.PROFILE):
        key = tag
        if tag == 'enc_tc':
            key = 'enc%d_tc' % (sum(1 for t in order if t == 'enc_tc') % 3)
        if tag == 'head_tc':
            key = 'head_last' if sum(1 for t in order if t == 'head_tc') % T == T - 1 else 'head_tc'
        order.append(tag)
        acc.setdefault(key, []).append(a.elapsed_time(b))
    _lib.PROFILE = None
    print('%-46s' % label, '  '.join('%s %.3f' % (k, sorted(v)[len(v) // 2]) for k, v in sorted(acc.items())
                                      if k in ('head_tc', 'head_last', 'enc0_tc', 'enc1_tc', 'enc2_tc')), flush=True)


for dbg, label in ((0, 'full kernels'), (1, 'epilogue: no math, no stores'), (5, 'epilogue: nothing (no TMEM loads either)'),
                   (8, 'epilogue: math, no stores'), (2, 'no MMAs (TMA + epilogue only)'), (10, 'no MMAs, no stores'),
                   (7, 'TMA only')):
    os.environ['ESSB_TC_DEBUG'] = str(dbg)
    run(label)
os.environ['ESSB_TC_DEBUG'] = '0'
for occ in ('1',):
    os.environ['ESSB_TC_OCC'] = occ
    print('(ESSB_TC_OCC is read once per process: run separately for occupancy A/B)')
