"""TEST INFRASTRUCTURE: build tests/emu/_build/libjsso_emu.so -- the product's kernel headers compiled by g++
(-DJSSO_EMU) against the SIMT emulator.  Used only by the CPU tests; the package never loads it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, '_build', 'libjsso_emu.so')
SRC = [os.path.join(HERE, 'emu_kernels.cpp'), os.path.join(ROOT, 'jaxsso_b200', 'csrc', 'jsso_symbolic.cpp')]
DEPS = SRC + [os.path.join(HERE, 'cuda_emu.h')] + \
    [os.path.join(ROOT, 'jaxsso_b200', 'csrc', f) for f in os.listdir(os.path.join(ROOT, 'jaxsso_b200', 'csrc'))]


def cuda_include():
    for c in (os.environ.get('CUDA_HOME'), '/usr/local/cuda'):
        if c and os.path.exists(os.path.join(c, 'include', 'cuda_runtime.h')):
            return os.path.join(c, 'include')
    raise RuntimeError('CUDA headers not found')


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ['g++', '-std=c++20', '-O1', '-g', '-fPIC', '-shared', '-pthread', '-DJSSO_EMU', '-w',
           # libjsso.so (loaded RTLD_GLOBAL by other tests) exports host stubs with the SAME mangled kernel names:
           # keep every reference inside this library
           '-fvisibility=hidden', '-fvisibility-inlines-hidden', '-Wl,-Bsymbolic',
           '-I' + cuda_include()] + SRC + ['-o', OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('emulator build failed:\n' + ' '.join(cmd) + '\n' + r.stderr[-6000:])
    return OUT


if __name__ == '__main__':
    print(build(force=True))
