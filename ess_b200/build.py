"""Builds ess_b200/libess_b200.so (in-tree) from ess_b200/csrc/*.cu with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting .so
travels to the GPU box with the repo snapshot.  `python -m ess_b200.build` or `build_library()`.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libess_b200.so')
SOURCES = ['api.cu', 'conv_fp32.cu', 'conv_tc.cu', 'wgrad_tc.cu', 'pointwise.cu', 'loss.cu', 'bn.cu', 'voxel.cu', 'pw_conv.cu', 'stem_conv.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I' + os.path.join(ROOT, 'include')]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(ROOT, 'include', 'ess_b200.h'), os.path.join(CSRC, 'common.cuh'),
               os.path.join(CSRC, 'tc_ptx.cuh')]
    nvcc = _nvcc()

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ['-c', s, '-o', o]
            if verbose:
                print(' '.join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return o

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))
