"""ms per distributed PCG iteration (fixed iteration count) -- torchrun ... scripts/dist_pcg_time.py SIZE ITERS"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from jaxsso_b200 import _native as nat, meshes, partition
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
nat.lib().jsso_set_device(local)
size = int(sys.argv[1]); iters = int(sys.argv[2])
md = meshes.plate(size)
owner = partition.rcb_owner(md.crds[:, :2], world)
lm = partition.local_mesh(md, owner, rank, world)
h = nat.Handle(lm.md.n_node, lm.md.cnct_quads, lm.md.cnct_beams, lm.md.known, device=local, n_row=lm.n_owned)
ids = [nat.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
h.set_halo(ids[0], rank, world, lm.peer_rank, lm.send_ptr, lm.send_idx, lm.recv_start, lm.recv_count)
if os.environ.get('JSSO_P2P', '1') != '0':
    hs = [None] * world
    dist.all_gather_object(hs, h.p2p_export())
    h.p2p_connect(hs, lm.remote_start)
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(lm.md.crds), D.from_host(lm.md.prop_quads), D.from_host(lm.md.prop_beams), D.from_host(lm.md.loads)
x = D((lm.md.ndof,))
for rep in range(2):
    h.assemble(crds, pq, pb, apply_bc=True)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = h.pcg(f, x, opts=nat.make_opts(rtol=1e-30, maxiter=iters, check_every=500), allow_noconv=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0:
    print(f'DIST_PCG world={world} p2p={os.environ.get("JSSO_P2P", "1")} size={size} iters={st.iterations} '
          f'ms_per_iter={1e3 * dt / iters:.4f} relres={st.relres:.3e}', flush=True)
dist.destroy_process_group()
