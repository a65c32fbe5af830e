import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from jaxsso_b200 import _native as nat, meshes, partition
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
nat.lib().jsso_set_device(local)
md = meshes.plate(16)
owner = partition.rcb_owner(md.crds[:, :2], world)
lm = partition.local_mesh(md, owner, rank, world)
h = nat.Handle(lm.md.n_node, lm.md.cnct_quads, lm.md.cnct_beams, lm.md.known, device=local, n_row=lm.n_owned)
ids = [nat.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
h.set_halo(ids[0], rank, world, lm.peer_rank, lm.send_ptr, lm.send_idx, lm.recv_start, lm.recv_count)
D = nat.DeviceArray
# 1. halo exchange of a vector whose value encodes the global dof id
x = np.full((lm.md.n_node, 6), -1.0)
x[:lm.n_owned] = 6.0 * lm.l2g[:lm.n_owned, None] + np.arange(6)
xd = D.from_host(x.ravel())
h.halo_exchange(xd)
got = xd.download().reshape(-1, 6)
exp = 6.0 * lm.l2g[:, None] + np.arange(6)
print(rank, 'halo ok', np.array_equal(got, exp), 'n_owned', lm.n_owned, 'n_node', lm.md.n_node, 'peers', lm.peer_rank, flush=True)
# 2. distributed spmv vs global
crds, pq, pb = D.from_host(lm.md.crds), D.from_host(lm.md.prop_quads), D.from_host(lm.md.prop_beams)
h.assemble(crds, pq, pb, apply_bc=True)
xg = np.random.default_rng(1).standard_normal(md.ndof)
xl = xg.reshape(-1, 6)[lm.l2g].ravel()
yd = D((6 * lm.n_owned,))
h.spmv(D.from_host(xl), yd)
if rank == 0:
    h1 = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=local)
    c1, q1, b1 = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams)
    h1.assemble(c1, q1, b1, apply_bc=True)
    y1 = D((md.ndof,)); h1.spmv(D.from_host(xg), y1)
    yref = y1.download().reshape(-1, 6)[lm.l2g[:lm.n_owned]].ravel()
    print('spmv err', np.abs(yd.download() - yref).max() / np.abs(yref).max(), flush=True)
# 3. pcg few iterations
f = D.from_host(lm.md.loads); u = D((lm.md.ndof,))
for mi in (1, 2, 5, 50):
    h.assemble(crds, pq, pb, apply_bc=True)
    try:
        st = h.pcg(f, u, opts=nat.make_opts(rtol=1e-10, maxiter=mi, check_every=1), allow_noconv=True)
        print(rank, mi, st.as_dict(), flush=True)
    except Exception as e:
        print(rank, mi, 'ERR', e, flush=True)
dist.destroy_process_group()
