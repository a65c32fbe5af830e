import numpy as np, sys
sys.path.insert(0, '.')
from jaxsso_b200 import _native as nat, meshes
from oracle import jaxsso_oracle as orc
from tests.conftest import to_oracle_mesh
md = meshes.plate(8)
h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
crds = nat.DeviceArray.from_host(md.crds); pq = nat.DeviceArray.from_host(md.prop_quads); pb = nat.DeviceArray.from_host(md.prop_beams)
f = nat.DeviceArray.from_host(md.loads)
h.assemble(crds, pq, pb, apply_bc=True)
rp, ci = h.pattern()
K = nat.bsr_to_scipy(rp, ci, h.values_host()).toarray()
print('K sym', np.abs(K-K.T).max()/np.abs(K).max(), 'min eig', np.linalg.eigvalsh(0.5*(K+K.T)).min())
x = nat.DeviceArray((md.ndof,))
for it in (1, 2, 3, 50):
    h.assemble(crds, pq, pb, apply_bc=True)
    try:
        st = h.pcg(f, x, opts=nat.make_opts(rtol=1e-10, maxiter=it, check_every=1), allow_noconv=True)
        print(it, st.as_dict())
    except Exception as e:
        print(it, 'ERR', e)
    Ks = nat.bsr_to_scipy(rp, ci, h.values_host()).toarray()
    # expected scaled matrix
    n = md.n_node
    W = np.zeros((6*n, 6*n))
    for r in range(n):
        D = K[6*r:6*r+6, 6*r:6*r+6]
        L = np.linalg.cholesky(0.5*(D+D.T))
        W[6*r:6*r+6, 6*r:6*r+6] = np.linalg.inv(L)
    Kexp = W @ K @ W.T
    print('  scaled matrix err', np.abs(Ks-Kexp).max(), 'diag', np.abs(np.diag(Ks)-1).max())
