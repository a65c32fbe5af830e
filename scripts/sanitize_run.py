"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family once."""
import sys
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
md = meshes.plate(12)
nid = np.arange(md.n_node).reshape(13, 13)
md.cnct_beams = np.stack([nid[6, :-1], nid[6, 1:]], 1).astype(np.int32)
md.prop_beams = np.tile([3.79e9, 3.79e9 / 2.6, 6.7e-5, 1.7e-5, 8.4e-5, 0.02], (12, 1))
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
h.mg_setup(max_coarse_nodes=8)
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
u, g = D((md.ndof,)), D.from_host(np.random.default_rng(0).standard_normal(md.ndof))
dc, dq, db = D((md.n_node, 3)), D((md.n_quad, 5)), D((md.n_beam, 6))
h.quad_ke(crds, pq); h.beam_ke(crds, pb)
for pre in ('block_jacobi', 'multigrid'):
    st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=1e-9, precond=pre, check_every=20))
    h.backward(crds, pq, pb, u, g, dc, dq, db, opts=nat.make_opts(rtol=1e-9, precond=pre, check_every=20))
    print(pre, st.as_dict())
val = h.value_and_grad_host(md.crds, md.prop_quads, md.prop_beams, md.loads)[0]
print('ok', val)
# both assembly paths (warp tasks by default above, chunked single kernel here) and the 3-kernel CG
import os
os.environ['JSSO_ASM_CHUNKED'] = '1'
h2 = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
del os.environ['JSSO_ASM_CHUNKED']
h2.assemble(crds, pq, pb, apply_bc=True)
y = D((md.ndof,))
h.assemble(crds, pq, pb, apply_bc=True)
h.spmv(u, y)
print('chunked vs tasks', float(np.abs(h.values_host() - h2.values_host()).max()))
