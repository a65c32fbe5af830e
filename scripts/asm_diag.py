import sys, ctypes
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
md = meshes.plate(1024)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
D = nat.DeviceArray
crds, pq, pb = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
L = nat.lib()
ev = [L.jsso_event_create() for _ in range(2)]
for mode, name in ((1, 'full'), (3, 'no staging'), (5, 'no compute (staging + zero store)'), (7, 'store only')):
    for _ in range(3):
        L.jsso_assemble(h.h, crds.ptr, pq.ptr, pb.ptr, mode, None)
    L.jsso_event_record(ev[0], None)
    for _ in range(10):
        L.jsso_assemble(h.h, crds.ptr, pq.ptr, pb.ptr, mode, None)
    L.jsso_event_record(ev[1], None)
    ms = ctypes.c_float(); L.jsso_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms))
    print(f'{name:40s} {ms.value/10:.3f} ms')
