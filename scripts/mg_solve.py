import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
N = int(sys.argv[1]); deg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
md = meshes.plate(N)
t0 = time.perf_counter(); h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0); t1 = time.perf_counter()
h.mg_setup(); t2 = time.perf_counter()
print('handle %.2fs mg symbolic %.2fs levels %s' % (t1 - t0, t2 - t1, h.mg_levels), flush=True)
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
u = D((md.ndof,))
for rep in range(2):
    t0 = time.perf_counter()
    try:
        st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=1e-8, precond='multigrid', cheb_degree=deg))
        print('rep', rep, 'forward %.3fs' % (time.perf_counter() - t0), st.as_dict(), flush=True)
    except Exception as e:
        print('ERR', e); break
# setup cost: one iteration only
t0 = time.perf_counter()
try:
    h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=1e-8, precond='multigrid', cheb_degree=deg, maxiter=1))
except Exception as e:
    pass
print('assemble + scaling + numeric mg setup + 1 iteration: %.3fs' % (time.perf_counter() - t0), flush=True)
