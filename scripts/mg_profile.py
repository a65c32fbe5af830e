"""A few multigrid-PCG iterations on one plate for ncu.  The numeric setup runs first (un-profiled); the profiled
range (cudaProfilerStart/Stop) holds `iters` PCG iterations of jsso_pcg on the already-built hierarchy:
  ncu --profile-from-start off --set full --clock-control none -o out python scripts/mg_profile.py 1024 3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsso_b200 import _native as nat, meshes
N = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
deg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
what = sys.argv[4] if len(sys.argv) > 4 else 'iterations'     # 'setup': profile assembly + scaling + numeric hierarchy instead
md = meshes.plate(N)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
h.mg_setup()
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
u = D((md.ndof,))
opts = nat.make_opts(rtol=1e-8, precond='multigrid', maxiter=iters, cheb_degree=deg)
try:
    h.forward(crds, pq, pb, f, u, opts=opts)          # assembly + scaling + numeric hierarchy + `iters` iterations
except nat.JssoError:
    pass
if what == 'setup':
    nat.lib().jsso_profiler_range(1)
    try:
        h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=1e-8, precond='multigrid', maxiter=1, cheb_degree=deg))
    except nat.JssoError:
        pass
    nat.lib().jsso_profiler_range(0)
    sys.exit(0)
nat.lib().jsso_profiler_range(1)
try:
    h.pcg(f, u, opts=opts)                            # same matrix, same hierarchy: only the PCG iterations
except nat.JssoError as e:
    print('stopped:', e)
nat.lib().jsso_profiler_range(0)
