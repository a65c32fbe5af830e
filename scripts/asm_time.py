"""Time the assembly (CUDA events) of the library selected by JSSO_LIB and check it against the chunked
single-kernel path of the same library:  python scripts/asm_time.py N"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
md = meshes.plate(N)
L = nat.lib()
D = nat.DeviceArray
crds, pq, pb = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
ev = [L.jsso_event_create() for _ in range(2)]
for _ in range(5):
    h.assemble(crds, pq, pb, apply_bc=True)
L.jsso_stream_sync(None)
reps = 30
L.jsso_event_record(ev[0], None)
for _ in range(reps):
    h.assemble(crds, pq, pb, apply_bc=True)
L.jsso_event_record(ev[1], None)
ms = ctypes.c_float()
L.jsso_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms))
v = h.values_host()
err = -1.0
if os.environ.get('JSSO_CHECK', '1') == '1' and N <= 512:
    os.environ['JSSO_ASM_CHUNKED'] = '1'
    h2 = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    h2.assemble(crds, pq, pb, apply_bc=True)
    v2 = h2.values_host()
    err = float(np.abs(v - v2).max() / np.abs(v2).max())
print(f'assemble {ms.value / reps:.4f} ms  relerr_vs_chunked {err:.2e}  checksum {float(np.abs(v).sum()):.10e}')
