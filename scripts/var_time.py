"""Assembly stage split (geometry / tasks) and multigrid solve time of the library selected by JSSO_LIB."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
md = meshes.plate(N)
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
h.profile(True)
g, t = [], []
for i in range(25):
    h.assemble(crds, pq, pb, apply_bc=True)
    a, b = h.profile_read()
    if i >= 5:
        g.append(a); t.append(b)
h.profile(False)
h.mg_setup()
u = D((md.ndof,))
best = 1e9
for rep in range(2):
    t0 = time.perf_counter()
    st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=1e-8, precond='multigrid', cheb_degree=1))
    best = min(best, time.perf_counter() - t0)
print(f'geometry {np.mean(g):.4f} ms  tasks {np.mean(t):.4f} ms  mg forward {best:.4f} s ({st.iterations} it)')
