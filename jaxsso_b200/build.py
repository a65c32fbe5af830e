"""Build libjsso.so (sm_100a) in-tree with nvcc.  No JIT cache, no torch."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libjsso.so')
SOURCES = ['jsso_api.cu', 'jsso_symbolic.cpp']
HEADERS = ['jsso_elem.cuh', 'jsso_assemble.cuh', 'jsso_solver.cuh', 'jsso_adjoint.cuh', 'jsso_multigrid.cuh',
           'jsso_symbolic.h',
           os.path.join('..', '..', 'include', 'jsso.h')]


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defs=(), out=None):
    """Compile every CUDA source for sm_100a into jaxsso_b200/libjsso.so (`defs`/`out`: tuning
    variants built beside it, selected at run time with JSSO_LIB=<path>)."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-O3']
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += ['-D' + d for d in defs]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ['-o', out or LIB, '-lcudart', '-ldl']   # NCCL is bound lazily with dlopen (see jsso_api.cu)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
    return out or LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
