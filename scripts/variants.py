"""Build tuning variants of libjsso.so (compile-time switches of the task assembly kernel) and time
them on the GPU box:   python scripts/variants.py build      (here, no GPU)
                       python scripts/variants.py run [N]    (on the box: one process per variant)"""
import os, subprocess, sys
sys.path.insert(0, '.')
VARIANTS = {
    'base':   [],
    'g64':    ['JSSO_G_QUADS=64'],
    'g128':   ['JSSO_G_QUADS=128'],
    'g64t256': ['JSSO_G_QUADS=64', 'JSSO_G_THREADS=256'],
    'g128t256': ['JSSO_G_QUADS=128', 'JSSO_G_THREADS=256'],
    'spmv5':  ['JSSO_SPMV_MINB=5'],
    'spmv6':  ['JSSO_SPMV_MINB=6'],
}
VDIR = os.path.join('jaxsso_b200', 'variants')


def build():
    from jaxsso_b200 import build as jb
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        out = os.path.abspath(os.path.join(VDIR, f'libjsso_{name}.so'))
        jb.build(force=True, defs=defs, out=out)
        print('built', out, defs, flush=True)


def run(N):
    for name in VARIANTS:
        env = dict(os.environ, JSSO_LIB=os.path.abspath(os.path.join(VDIR, f'libjsso_{name}.so')))
        r = subprocess.run([sys.executable, os.environ.get('JSSO_VARIANT_SCRIPT', 'scripts/asm_time.py'), str(N)], env=env, capture_output=True, text=True)
        print(f'{name:22s}', (r.stdout.strip().splitlines() or ['?'])[-1], r.stderr.strip()[-300:], flush=True)


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
