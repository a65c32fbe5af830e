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


def _match_back(src, i):
    """index of the start of the kernel name expression that ends at src[i] (exclusive): identifier + optional <...>"""
    j = i
    while j > 0 and src[j - 1].isspace():
        j -= 1
    if src[j - 1] == '>':            # template arguments
        depth = 0
        while j > 0:
            j -= 1
            if src[j] == '>':
                depth += 1
            elif src[j] == '<':
                depth -= 1
                if depth == 0:
                    break
    while j > 0 and (src[j - 1].isalnum() or src[j - 1] in '_:'):
        j -= 1
    return j


def _split_top(s):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur.strip())
            cur = ''
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    """`kernel<<<grid, block, smem, stream>>>(args);`  ->  `emu::launch(grid, block, smem, [=] { kernel(args); });`"""
    out, pos, n = [], 0, 0
    while True:
        i = src.find('<<<', pos)
        if i < 0:
            break
        k = src.index('>>>', i)
        cfg = _split_top(src[i + 3:k])
        start = _match_back(src, i)
        name = src[start:i].strip()
        a = src.index('(', k)
        depth, e = 0, a
        while True:
            if src[e] == '(':
                depth += 1
            elif src[e] == ')':
                depth -= 1
                if depth == 0:
                    break
            e += 1
        args = src[a + 1:e]
        assert src[e + 1] == ';', src[e:e + 20]
        smem = cfg[2] if len(cfg) > 2 else '0'
        out.append(src[pos:start])
        # the arguments are evaluated ONCE, at launch, and copied (CUDA semantics): an argument expression with a side
        # effect inside the per-thread closure would run once per emulated thread
        out.append(f'{{ auto emu_args_ = std::make_tuple({args}); emu::launch({cfg[0]}, {cfg[1]}, {smem}, '
                   f'[=] {{ std::apply([](auto... a_) {{ {name}(a_...); }}, emu_args_); }}); }}')
        pos = e + 2
        n += 1
    out.append(src[pos:])
    return ''.join(out), n


API_OUT = os.path.join(HERE, '_build', 'libjsso_emuapi.so')
API_ASAN_OUT = os.path.join(HERE, '_build', 'libjsso_emuapi_asan.so')


def asan_runtime():
    """Path of libasan.so for LD_PRELOAD (the Python interpreter itself is not instrumented)."""
    r = subprocess.run(['g++', '-print-file-name=libasan.so'], capture_output=True, text=True)
    p = r.stdout.strip()
    return os.path.realpath(p) if os.path.sep in p and os.path.exists(p) else None


def build_api(force=False, asan=False):
    """The product's host driver (csrc/jsso_api.cu) on the emulator: same C ABI as libjsso.so, for the CPU tests
    only (select it in a SUBPROCESS with JSSO_LIB=<path>; the package itself never looks for it)."""
    api = os.path.join(ROOT, 'jaxsso_b200', 'csrc', 'jsso_api.cu')
    deps = DEPS + [os.path.join(HERE, 'cuda_emu_rt.h'), os.path.join(ROOT, 'include', 'jsso.h'), os.path.abspath(__file__)]
    target = API_ASAN_OUT if asan else API_OUT
    if not force and os.path.exists(target) and all(os.path.getmtime(d) <= os.path.getmtime(target) for d in deps):
        return target
    os.makedirs(os.path.dirname(target), exist_ok=True)
    src, n = rewrite_launches(open(api).read())
    assert n > 50 and '<<<' not in src
    csrc = os.path.join(ROOT, 'jaxsso_b200', 'csrc')
    src = src.replace('#include "../../include/jsso.h"', '#pragma GCC visibility push(default)\n#include "%s"\n#pragma GCC visibility pop'
                      % os.path.join(ROOT, 'include', 'jsso.h'))
    for hname in ('jsso_adjoint.cuh', 'jsso_assemble.cuh', 'jsso_solver.cuh', 'jsso_multigrid.cuh', 'jsso_symbolic.h'):
        src = src.replace('#include "%s"' % hname, '#include "%s"' % os.path.join(csrc, hname))
    gen = os.path.join(HERE, '_build', 'jsso_api_emu_asan.cpp' if asan else 'jsso_api_emu.cpp')
    with open(gen, 'w') as f:
        f.write('// GENERATED by tests/emu/build_emu.py from jaxsso_b200/csrc/jsso_api.cu (kernel launches rewritten)\n')
        f.write('#include "%s"\n' % os.path.join(HERE, 'cuda_emu_rt.h'))
        f.write(src)
    nccl_inc = []
    for c in ('/usr/include', '/usr/local/cuda/include'):
        if os.path.exists(os.path.join(c, 'nccl.h')):
            nccl_inc = ['-I' + c]
            break
    if not nccl_inc:
        import glob
        hits = glob.glob('/opt/**/nccl.h', recursive=True) + glob.glob(os.path.join(os.path.dirname(os.__file__), 'site-packages', 'nvidia', 'nccl', 'include', 'nccl.h'))
        if hits:
            nccl_inc = ['-I' + os.path.dirname(hits[0])]
    cmd = ['g++', '-std=c++20', '-O1', '-g', '-fPIC', '-shared', '-pthread', '-DJSSO_EMU', '-w',
           '-fvisibility=hidden', '-fvisibility-inlines-hidden', '-Wl,-Bsymbolic', '-I' + cuda_include()] + nccl_inc + \
          [gen, os.path.join(csrc, 'jsso_symbolic.cpp'), '-o', target, '-ldl']
    if asan:   # every "device" buffer is a heap block: out-of-bounds kernel accesses become AddressSanitizer reports
        cmd[1:1] = ['-fsanitize=address', '-fno-omit-frame-pointer']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('emulated API build failed:\n' + ' '.join(cmd) + '\n' + r.stderr[-8000:])
    return target


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
    print(build_api(force=True))
