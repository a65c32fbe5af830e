"""Run each hot kernel a few times on one plate (for ncu): python scripts/prof_kernels.py SIZE [solve_iters]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
sys.path.insert(0, '.')
from bench import synthetic_state
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
md = meshes.plate(N)
u, lam = synthetic_state(md)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
D = nat.DeviceArray
crds, pq, pb = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
u_d, lam_d, f_d = D.from_host(u), D.from_host(lam), D.from_host(md.loads)
dc, dq = D((md.n_node, 3)), D((md.n_quad, 5))
for _ in range(3):
    h.assemble(crds, pq, pb, apply_bc=True)
    h.adjoint(crds, pq, pb, u_d, lam_d, dc, dq, None)
x = D((md.ndof,))
st = h.pcg(f_d, x, opts=nat.make_opts(rtol=1e-30, maxiter=iters, check_every=iters), allow_noconv=True)
nat.lib().jsso_stream_sync(None)
print('done', st.as_dict())
