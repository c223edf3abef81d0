"""Build libmft_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmft_b200.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']
# kernels.cu carries the bit-exact chain/select/lookup arithmetic: no FMA contraction there.
SOURCES = [('conv_tc.cu', []), ('kernels.cu', ['--fmad=false']), ('engine.cu', [])]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build_library(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', 'nvcc')
    objdir = os.path.join(os.path.dirname(HERE), 'build')
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'mft_b200.h'))
    objs, relink = [], force or not os.path.exists(LIB)
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [nvcc] + ARCH + COMMON + extra + ['-c', s, '-o', o]
            if verbose:
                print(' '.join(cmd))
            subprocess.check_call(cmd)
            relink = True
        objs.append(o)
    if relink:
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ARCH
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build_library(verbose=True))
