import numpy as np, sys
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
from oracle import jaxsso_oracle as orc
from tests.conftest import to_oracle_mesh
md = meshes.mannheim_quad(dict(np.load('tests/golden/mannheim_quad.npz')))
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
crds = nat.DeviceArray.from_host(md.crds); pq = nat.DeviceArray.from_host(md.prop_quads); pb = nat.DeviceArray.from_host(md.prop_beams)
f = nat.DeviceArray.from_host(md.loads)
u = nat.DeviceArray((md.ndof,))
uref = orc.solve_refined(to_oracle_mesh(md))
for rtol in (1e-8, 1e-10, 1e-11, 1e-12):
    try:
        st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=rtol))
        print(rtol, st.as_dict(), 'u err', np.linalg.norm(u.download()-uref)/np.linalg.norm(uref))
    except Exception as e:
        print(rtol, 'ERR', e, 'u err', np.linalg.norm(u.download()-uref)/np.linalg.norm(uref))
