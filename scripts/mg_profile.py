"""A few multigrid-PCG iterations on one plate, for an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/mg_profile.py 1024 3"""
import sys
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
N = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
md = meshes.plate(N)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
h.mg_setup()
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
u = D((md.ndof,))
try:
    st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=1e-8, precond='multigrid', maxiter=iters, cheb_degree=int(sys.argv[3]) if len(sys.argv) > 3 else 1))
    print(st.as_dict())
except Exception as e:
    print('stopped:', e)
