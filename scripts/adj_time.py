"""Time the adjoint stage (CUDA events) of the library selected by JSSO_LIB:  python scripts/adj_time.py N"""
import ctypes, sys
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
from bench import synthetic_state
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
md = meshes.plate(N)
u, lam = synthetic_state(md)
L = nat.lib()
D = nat.DeviceArray
crds, pq, pb = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
u_d, lam_d = D.from_host(u), D.from_host(lam)
dc, dq = D((md.n_node, 3)), D((md.n_quad, 5))
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
ev = [L.jsso_event_create() for _ in range(2)]
res = []
for want_prop in (True, False):
    for _ in range(5):
        h.adjoint(crds, pq, pb, u_d, lam_d, dc, dq if want_prop else None, None)
    L.jsso_stream_sync(None)
    L.jsso_event_record(ev[0], None)
    for _ in range(30):
        h.adjoint(crds, pq, pb, u_d, lam_d, dc, dq if want_prop else None, None)
    L.jsso_event_record(ev[1], None)
    ms = ctypes.c_float()
    L.jsso_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms))
    res.append(ms.value / 30)
g = dc.download()
print(f'adjoint with d_prop {res[0]:.4f} ms, without {res[1]:.4f} ms  checksum {float(np.abs(g).sum()):.12e}')
