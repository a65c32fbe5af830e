"""Distributed (one process per GPU) forward + backward on one partitioned mesh, compared with the
single-GPU result computed on rank 0.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/dist_check.py [SIZE]
Prints one line `DIST_CHECK {...json...}` on rank 0 and exits non-zero on mismatch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from jaxsso_b200 import _native as nat
from jaxsso_b200 import meshes, partition

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nat.lib().jsso_set_device(local)
md = meshes.plate(size)
owner = partition.rcb_owner(md.crds[:, :2], world)
lm = partition.local_mesh(md, owner, rank, world)
h = nat.Handle(lm.md.n_node, lm.md.cnct_quads, lm.md.cnct_beams, lm.md.known, device=local, n_row=lm.n_owned)
ids = [nat.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
h.set_halo(ids[0], rank, world, lm.peer_rank, lm.send_ptr, lm.send_idx, lm.recv_start, lm.recv_count)
use_p2p = os.environ.get('JSSO_P2P', '1') != '0'
if use_p2p:
    hs = [None] * world
    dist.all_gather_object(hs, h.p2p_export())
    h.p2p_connect(hs, lm.remote_start)
D = nat.DeviceArray
crds, pq, pb, f = D.from_host(lm.md.crds), D.from_host(lm.md.prop_quads), D.from_host(lm.md.prop_beams), D.from_host(lm.md.loads)
u = D((lm.md.ndof,))
dc, dq = D((lm.md.n_node, 3)), D((lm.md.n_quad, 5))
opts = nat.make_opts(rtol=1e-11, compliance=True)
fs = h.forward(crds, pq, pb, f, u, opts=opts)
h.backward(crds, pq, pb, u, None, dc, dq, None, opts=opts)
ul = u.download().reshape(-1, 6)[:lm.n_owned]
dcl = dc.download()[:lm.n_owned]
# gather owned results on rank 0
gu = [None] * world
dist.all_gather_object(gu, (lm.l2g[:lm.n_owned], ul, dcl, lm.quad_ids, dq.download(), fs.iterations, fs.relres))
ok = True
if rank == 0:
    U = np.zeros((md.n_node, 6)); G = np.zeros((md.n_node, 3)); Q = np.zeros((md.n_quad, 5))
    for ids_, u_, g_, qi, q_, it, rr in gu:
        U[ids_] = u_; G[ids_] = g_; Q[qi] = q_       # ghost-element copies of d_prop agree, last writer wins
    h1 = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=local)
    val, u1, dc1, dq1, _, fs1, _ = h1.value_and_grad_host(md.crds, md.prop_quads, md.prop_beams, md.loads,
                                                         opts=nat.make_opts(rtol=1e-11))
    eu = np.linalg.norm(U.ravel() - u1) / np.linalg.norm(u1)
    eg = np.abs(G - dc1).max() / np.abs(dc1).max()
    eq = np.abs(Q - dq1).max() / np.abs(dq1).max()
    res = {'world': world, 'size': size, 'p2p': use_p2p, 'u_err': eu, 'grad_err': eg, 'dprop_err': eq,
           'iters_dist': gu[0][5], 'iters_single': fs1.iterations, 'relres_dist': gu[0][6]}
    print('DIST_CHECK', json.dumps(res))
    ok = eu < 1e-8 and eg < 1e-7 and eq < 1e-7
flag = torch.tensor([1 if ok else 0], device='cuda')
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
