"""The C-ABI library loads without a GPU and exports every symbol include/ess_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'ess_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(essb_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_header_symbols():
    from ess_b200 import _lib, build
    build.build_library()
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), 'missing export %s' % name
    assert set(declared) == set(_lib.SIGNATURES.keys()), set(declared) ^ set(_lib.SIGNATURES.keys())
    assert lib.essb_version() == 100
    assert lib.essb_build_arch() == b'sm_100a'


def test_ctypes_struct_sizes_match_header():
    """sizeof of the ctypes mirrors == sizeof of the C structs (compiled with gcc from the header)."""
    import ctypes
    import subprocess
    import tempfile
    from ess_b200 import _lib
    prog = ('#include <stdio.h>\n#include "ess_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(essb_src),'
            ' sizeof(essb_conv), sizeof(essb_wgrad), sizeof(essb_tc_view), sizeof(essb_conv_tc)); '
            'printf("%zu %zu\\n", sizeof(essb_wgrad_tc), sizeof(essb_radam_multi)); return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 's.c')
        open(c, 'w').write(prog)
        exe = os.path.join(d, 's')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mine = [ctypes.sizeof(t) for t in (_lib.Src, _lib.Conv, _lib.Wgrad, _lib.TcView, _lib.ConvTc, _lib.WgradTc,
                                      _lib.RadamMulti)]
    assert sizes == mine, (sizes, mine)


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    import torch
    import ess_b200
    m = ess_b200.E2VIDRecurrent(dict(num_bins=2, num_encoders=3, base_num_channels=4, norm='BN',
                                     use_upsample_conv=False), mode='fp32')
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 16, 16), None)
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=3, ignore_index=255)
    with pytest.raises(RuntimeError):
        crit(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'ess_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            assert 'oracle' not in open(os.path.join(pkg, fn)).read(), fn


def _prototypes():
    """name -> list of C parameter type strings, parsed from the header's function declarations."""
    src = open(os.path.join(ROOT, 'include', 'ess_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(?:int|int64_t|const char\s*\*)\s+(essb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = [a.strip() for a in m.group(2).replace('\n', ' ').split(',')]
        protos[m.group(1)] = [] if args in ([''], ['void']) else args
    return protos


def test_ctypes_signatures_match_header_prototypes():
    """Every ctypes signature has the header's parameter count and the same kind per parameter (pointer /
    32-bit int / 64-bit int / float): a drifted hand-written mirror would corrupt the call frame silently."""
    import ctypes as C
    from ess_b200 import _lib
    protos = _prototypes()
    assert set(protos) == set(_lib.SIGNATURES), set(protos) ^ set(_lib.SIGNATURES)

    def kind_c(t):
        t = re.sub(r'\b[a-zA-Z_][a-zA-Z0-9_]*\s*$', '', t).strip() if not t.endswith('*') else t   # drop the name
        if '*' in t:
            return 'ptr'
        if 'int64_t' in t or 'long long' in t:
            return 'i64'
        if 'float' in t:
            return 'f32'
        if 'double' in t:
            return 'f64'
        return 'i32'

    def kind_py(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return 'ptr'
        return {C.c_int: 'i32', C.c_int32: 'i32', C.c_int64: 'i64', C.c_float: 'f32', C.c_double: 'f64'}[t]

    for name, (res, args) in _lib.SIGNATURES.items():
        cargs = protos[name]
        assert len(cargs) == len(args), '%s: header has %d parameters, ctypes mirror %d' % (name, len(cargs), len(args))
        for i, (ca, pa) in enumerate(zip(cargs, args)):
            assert kind_c(ca) == kind_py(pa), '%s: parameter %d is `%s` in the header but %s in _lib.py' % (name, i, ca, pa)
