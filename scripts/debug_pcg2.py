import numpy as np, sys
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
for name, md in (('plate8', meshes.plate(8)), ('barrel', meshes.barrel_arch())):
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    crds = nat.DeviceArray.from_host(md.crds); pq = nat.DeviceArray.from_host(md.prop_quads); pb = nat.DeviceArray.from_host(md.prop_beams)
    f = nat.DeviceArray.from_host(md.loads)
    x = nat.DeviceArray((md.ndof,))
    for ce, mi in ((1, 200), (7, 200), (50, 200), (50, 100000)):
        h.assemble(crds, pq, pb, apply_bc=True)
        try:
            st = h.pcg(f, x, opts=nat.make_opts(rtol=1e-12, maxiter=mi, check_every=ce), allow_noconv=True)
            print(name, ce, mi, st.as_dict())
        except Exception as e:
            print(name, ce, mi, 'ERR', e)
