"""Assembly only, a few launches on one plate (for ncu): python scripts/prof_asm.py SIZE"""
import sys
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
md = meshes.plate(N)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
D = nat.DeviceArray
crds, pq, pb = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
for _ in range(3):
    h.assemble(crds, pq, pb, apply_bc=True)
nat.lib().jsso_stream_sync(None)
print('done')
