"""Row-range distributed multigrid solve (jsso_mg_set_dist) on N GPUs against the single-GPU multigrid solve of
the same system.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/dist_mg_check.py [SIZE] [MIN_DIST_NODES] [CHEB_DEGREE] [p2p]
Every rank builds the whole renumbered mesh, assembles it, and solves K u = f with the V-cycle PCG distributed by
row ranges; rank 0 then repeats the solve on an undistributed handle.  Prints `DIST_MG_CHECK {...json...}` on
rank 0 and exits non-zero on mismatch."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from jaxsso_b200 import _native as nat
from jaxsso_b200 import dist_multigrid as dmg
from jaxsso_b200 import meshes, partition

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
min_dist = int(sys.argv[2]) if len(sys.argv) > 2 else 500
deg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
use_p2p = 'p2p' in sys.argv[4:]    # exchanges over peer memory (jsso_mg_p2p_connect) instead of NCCL
dist_setup = 'setup' in sys.argv[4:]   # numeric setup distributed too (jsso_mg_set_dist_setup)
rtol = 1e-10
nat.lib().jsso_set_device(local)
md0 = meshes.plate(size)
owner = partition.rcb_owner(md0.crds[:, :2], world)
perm, bounds = dmg.owner_permutation(owner, world)
md = dmg.renumber_mesh(md0, perm)
D = nat.DeviceArray


def solve(distributed):
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=local)
    levels = h.mg_setup(max_coarse_nodes=100)
    info = None
    if distributed:
        rp, ci = h.pattern()
        plan = dmg.build_plan(rp, ci, levels, bounds, min_dist_nodes=min_dist)
        ids = [nat.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        h.mg_set_dist(ids[0], rank, world, plan)
        if dist_setup:
            h.mg_set_dist_setup(rp, ci, levels, plan, rank)
        if use_p2p:
            def allgather(obj):
                box = [None] * world
                dist.all_gather_object(box, obj)
                return box
            h.mg_p2p_connect(plan, allgather)
        info = dmg.plan_summary(plan)
    crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
    u = D((md.ndof,))
    opts = nat.make_opts(rtol=rtol, precond='multigrid', cheb_degree=deg)
    st = h.forward(crds, pq, pb, f, u, opts=opts)       # includes the numeric multigrid setup
    nat.lib().jsso_stream_sync(None)
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    st = h.forward(crds, pq, pb, f, u, opts=opts)
    nat.lib().jsso_stream_sync(None)
    dt = time.perf_counter() - t0
    cnt = h.mg_dist_counters() if distributed else (0, 0)
    out = u.download()
    h.close()
    return out, st, dt, info, cnt


ud, std, dtd, info, cnt = solve(True)
ok = True
# every rank must hold the same whole solution
chk = torch.tensor([float(np.abs(ud).sum()), float(ud @ ud)], dtype=torch.float64, device='cuda')
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool(torch.equal(lo, hi))
if rank == 0:
    us, sts, dts, _, _ = solve(False)
    eu = np.linalg.norm(ud - us) / np.linalg.norm(us)
    # and against the original numbering solved without multigrid
    h0 = nat.Handle(md0.n_node, md0.cnct_quads, md0.cnct_beams, md0.known, device=local)
    _, u0, *_ = h0.value_and_grad_host(md0.crds, md0.prop_quads, md0.prop_beams, md0.loads,
                                       opts=nat.make_opts(rtol=1e-11, precond='block_jacobi'))
    e0 = np.linalg.norm(ud.reshape(-1, 6) - u0.reshape(-1, 6)[perm]) / np.linalg.norm(u0)
    res = {'world': world, 'size': size, 'cheb_degree': deg, 'peer_memory': use_p2p, 'distributed_setup': dist_setup, 'u_err_vs_single_mg': eu, 'u_err_vs_block_jacobi': e0,
           'iters_dist': std.iterations, 'iters_single': sts.iterations, 'relres_dist': std.relres,
           'seconds_dist': dtd, 'seconds_single': dts, 'same_on_all_ranks': same, 'plan': info,
           'exchanges': cnt[0], 'allreduces': cnt[1]}
    print('DIST_MG_CHECK', json.dumps(res))
    ok = (eu < 1e-8 and e0 < 1e-7 and same and std.converged == 1 and
          abs(std.iterations - sts.iterations) <= 2)
flag = torch.tensor([1 if ok else 0], device='cuda')
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
