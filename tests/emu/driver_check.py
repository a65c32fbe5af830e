"""TEST INFRASTRUCTURE: runs the package's normal Python -> C ABI path against the EMULATED library
(JSSO_LIB=tests/emu/_build/libjsso_emuapi.so: the product's host driver + kernels compiled by g++ on the SIMT
emulator).  Launched as a subprocess by tests/test_emu_driver.py; prints one JSON line.

  driver_check.py grad  SIZE                         value + gradient of a mixed quad/beam plate vs the oracle
  driver_check.py e2e   SIZE                         jsso_assemble_adjoint_host (JSSO_E2E_CHUNKS from the environment)
  driver_check.py mg    SIZE DEG                     multigrid PCG vs block-Jacobi CG vs the oracle
  driver_check.py part  WORLD SIZE                   partitioned handles + distributed block-Jacobi CG (hardware-measured path)
  driver_check.py benchleg WORLD SIZE MIN_DIST [rcb] bench.py's N > 1 gradient evaluation (distributed_solve_handle) on rank threads
  driver_check.py dist  WORLD SIZE MIN_DIST DEG      row-range distributed multigrid solve on WORLD rank THREADS
                                                     (fake NCCL between them) vs the undistributed solve
"""
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from jaxsso_b200 import _native as nat                       # noqa: E402
from jaxsso_b200 import dist_multigrid as dmg               # noqa: E402
from jaxsso_b200 import meshes, partition                    # noqa: E402
from oracle import jaxsso_oracle as orc                      # noqa: E402

assert 'emuapi' in nat.LIB_PATH, 'this script is for the emulated library only'
D = nat.DeviceArray


def omesh(md):
    return orc.Mesh(md.crds, md.cnct_quads, md.prop_quads, md.cnct_beams, md.prop_beams, md.known, md.loads)


def mixed(n):
    md = meshes.plate(n)
    nid = np.arange(md.n_node).reshape(n + 1, n + 1)
    md.cnct_beams = np.stack([nid[n // 2, :-1], nid[n // 2, 1:]], 1).astype(np.int32)
    md.prop_beams = np.tile([3.79e9, 3.79e9 / 2.6, 6.7e-5, 1.7e-5, 8.4e-5, 0.02], (n, 1))
    return md


def solve(md, precond, deg=1, rtol=1e-10, dist=None, max_coarse=int(os.environ.get('EMU_MAX_COARSE', '8'))):
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
    info, cnt = None, (0, 0)
    if precond == 'multigrid':
        levels = h.mg_setup(max_coarse_nodes=max_coarse)
        if dist is not None:
            nid, rank, world, bounds, min_dist = dist[:5]
            rp, ci = h.pattern()
            plan = dmg.build_plan(rp, ci, levels, bounds, min_dist_nodes=min_dist)
            h.mg_set_dist(nid, rank, world, plan)
            if os.environ.get('EMU_DIST_SETUP', '0') == '1':   # numeric setup distributed as well
                h.mg_set_dist_setup(rp, ci, levels, plan, rank, _drop_ghost=os.environ.get('EMU_DROP_GHOST', '0') == '1')
            if len(dist) > 5 and dist[5] is not None:      # peer-memory exchange: gather the IPC bytes between the threads
                box, bar = dist[5]

                def allgather(b):
                    box[rank] = b
                    bar.wait()
                    out_ = list(box)
                    bar.wait()
                    return out_
                h.mg_p2p_connect(plan, allgather)
            info = dmg.plan_summary(plan)
    crds, pq, pb, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
    u = D((md.ndof,))
    st = h.forward(crds, pq, pb, f, u, opts=nat.make_opts(rtol=rtol, precond=precond, cheb_degree=deg))
    out = u.download()
    if dist is not None:
        cnt = h.mg_dist_counters() + (h.mg_dist_p2p,)
    h.close()
    return out, st.iterations, bool(st.converged), st.relres, info, cnt


def main():
    mode = sys.argv[1]
    if os.environ.get('EMU_DEBUG_HANG'):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ['EMU_DEBUG_HANG']), exit=True)
    if mode == 'grad':
        md = mixed(int(sys.argv[2]))
        h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
        val, u, dc, dq, db, fs, bs = h.value_and_grad_host(md.crds, md.prop_quads, md.prop_beams, md.loads,
                                                           opts=nat.make_opts(rtol=1e-11))
        rv, ru, rl, rdc, rdq, rdb = orc.value_and_grad(omesh(md))
        res = {'u_err': float(np.linalg.norm(u - ru) / np.linalg.norm(ru)), 'c_err': float(abs(val - rv) / abs(rv)),
               'g_err': float(np.abs(dc - rdc).max() / np.abs(rdc).max()),
               'dq_err': float(np.abs(dq - rdq).max() / np.abs(rdq).max()),
               'db_err': float(np.abs(db - rdb).max() / np.abs(rdb).max()),
               'iterations': fs.iterations, 'launches': int(nat.lib().jsso_launch_count())}
    elif mode == 'e2e':
        md = mixed(int(sys.argv[2]))
        rng = np.random.default_rng(5)
        u, lam = rng.standard_normal(md.ndof), rng.standard_normal(md.ndof)
        h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
        dc, dq, db = h.assemble_adjoint_host(md.crds, md.prop_quads, md.prop_beams, u, lam)
        K = h.values_host()
        rdc, rdq, rdb = orc.element_sensitivity(omesh(md), u, lam)
        res = {'chunks': os.environ.get('JSSO_E2E_CHUNKS', '1'),
               'g_err': float(np.abs(dc - rdc).max() / np.abs(rdc).max()),
               'dq_err': float(np.abs(dq - rdq).max() / np.abs(rdq).max()),
               'db_err': float(np.abs(db - rdb).max() / np.abs(rdb).max()),
               'sum': [float(dc.sum()), float(dq.sum()), float(db.sum()), float(np.abs(K).sum())],
               'hash': [hash(dc.tobytes()), hash(dq.tobytes()), hash(db.tobytes())]}
    elif mode == 'afterpcg':
        # jsso_spmv / jsso_get_values_host after a solve (the stored values are then the scaled W K W^T)
        md = mixed(int(sys.argv[2]))
        h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=0)
        c, q, b, f = D.from_host(md.crds), D.from_host(md.prop_quads), D.from_host(md.prop_beams), D.from_host(md.loads)
        h.assemble(c, q, b, apply_bc=True)
        K0 = h.values_host().copy()
        x = np.random.default_rng(4).standard_normal(md.ndof)
        xd, yd, ud = D.from_host(x), D((md.ndof,)), D((md.ndof,))
        h.spmv(xd, yd)
        y0 = yd.download()
        h.pcg(f, ud, opts=nat.make_opts(rtol=1e-10))
        K1 = h.values_host()
        h.spmv(xd, yd)
        res = {'K_err': float(np.abs(K1 - K0).max() / np.abs(K0).max()),
               'y_err': float(np.abs(yd.download() - y0).max() / np.abs(y0).max())}
    elif mode == 'mg':
        md, deg = meshes.plate(int(sys.argv[2])), int(sys.argv[3])
        l0 = int(nat.lib().jsso_launch_count())
        um, itm, cm, rm, _, _ = solve(md, 'multigrid', deg)
        mg_launches = int(nat.lib().jsso_launch_count()) - l0
        ub, itb, cb, rb, _, _ = solve(md, 'block_jacobi')
        uref = orc.solve_refined(omesh(md))
        res = {'mg_err': float(np.linalg.norm(um - uref) / np.linalg.norm(uref)),
               'bj_err': float(np.linalg.norm(ub - uref) / np.linalg.norm(uref)),
               'mg_iters': itm, 'bj_iters': itb, 'mg_converged': cm, 'bj_converged': cb, 'fp16': os.environ.get('JSSO_MG_FP16', '0'),
               'mg_launches': mg_launches}
    elif mode == 'dist':
        world, size, min_dist, deg = (int(a) for a in sys.argv[2:6])
        p2p = (([None] * world, threading.Barrier(world)) if (len(sys.argv) > 6 and sys.argv[6] == 'p2p') else None)
        md0 = meshes.plate(size)
        owner = partition.rcb_owner(md0.crds[:, :2], world)
        perm, bounds = dmg.owner_permutation(owner, world)
        md = dmg.renumber_mesh(md0, perm)
        nid = nat.nccl_unique_id()
        out = [None] * world

        def worker(rank):
            try:
                out[rank] = solve(md, 'multigrid', deg, dist=(nid, rank, world, bounds, min_dist, p2p))
            except BaseException as e:      # the other ranks would wait for this one forever
                import traceback
                print(f'EMU_RANK_ERROR rank {rank}: {e}', flush=True)
                traceback.print_exc()
                os._exit(3)

        th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        us, its, cs, rs, _, _ = solve(md, 'multigrid', deg)
        uref = orc.solve_refined(omesh(md0))
        res = {'world': world, 'iters_dist': [o[1] for o in out], 'iters_single': its,
               'converged': all(o[2] for o in out),
               'err_vs_single': float(max(np.linalg.norm(o[0] - us) for o in out) / np.linalg.norm(us)),
               'identical_on_all_ranks': all(np.array_equal(o[0], out[0][0]) for o in out),
               'err_vs_oracle': float(np.linalg.norm(out[0][0].reshape(-1, 6) - uref.reshape(-1, 6)[perm]) / np.linalg.norm(uref)),
               'plan': out[0][4], 'exchanges': out[0][5][0], 'allreduces': out[0][5][1], 'peer_memory': bool(out[0][5][2])}
    elif mode == 'part':
        # the partitioned path that IS measured on hardware (scripts/dist_check.py: set_halo + distributed block-Jacobi
        # CG over NCCL send/recv, partitioned adjoint) on rank threads: calibrates the fake NCCL against a known-good path
        world, size = int(sys.argv[2]), int(sys.argv[3])
        use_p2p = len(sys.argv) > 4 and sys.argv[4] == 'p2p'
        gmd = meshes.plate(size)
        owner = partition.rcb_owner(gmd.crds[:, :2], world)
        nid = nat.nccl_unique_id()
        out = [None] * world
        handles = [None] * world
        bar = threading.Barrier(world)
        opts = nat.make_opts(rtol=1e-11, compliance=True, precond='block_jacobi')

        def worker(rank):
            lm = partition.local_mesh(gmd, owner, rank, world)
            h = nat.Handle(lm.md.n_node, lm.md.cnct_quads, lm.md.cnct_beams, lm.md.known, device=0, n_row=lm.n_owned)
            h.set_halo(nid, rank, world, lm.peer_rank, lm.send_ptr, lm.send_idx, lm.recv_start, lm.recv_count)
            if use_p2p:      # peer-memory path: halo pushes and mailbox all-reduces inside the CG kernels
                handles[rank] = h.p2p_export()
                bar.wait()
                h.p2p_connect(list(handles), lm.remote_start)
                bar.wait()
            crds, pq, pb, f = (D.from_host(a) for a in (lm.md.crds, lm.md.prop_quads, lm.md.prop_beams, lm.md.loads))
            u, dc, dq = D((lm.md.ndof,)), D((lm.md.n_node, 3)), D((lm.md.n_quad, 5))
            fs = h.forward(crds, pq, pb, f, u, opts=opts)
            h.backward(crds, pq, pb, u, None, dc, dq, None, opts=opts)
            out[rank] = (lm.l2g[:lm.n_owned], u.download().reshape(-1, 6)[:lm.n_owned], dc.download()[:lm.n_owned],
                         lm.quad_ids, dq.download(), fs.iterations)

        th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        U, G, Q = np.zeros((gmd.n_node, 6)), np.zeros((gmd.n_node, 3)), np.zeros((gmd.n_quad, 5))
        for ids, u_, g_, qi, q_, it in out:
            U[ids] = u_; G[ids] = g_; Q[qi] = q_
        rv, ru, rl, rdc, rdq, rdb = orc.value_and_grad(omesh(gmd))
        res = {'u_err': float(np.linalg.norm(U.ravel() - ru) / np.linalg.norm(ru)),
               'g_err': float(np.abs(G - rdc).max() / np.abs(rdc).max()),
               'dq_err': float(np.abs(Q - rdq).max() / np.abs(rdq).max()), 'iterations': [o[5] for o in out]}
    elif mode == 'benchleg':
        # bench.py's N > 1 gradient evaluation on rank threads: partitioned handles (RCB) for the adjoint, the whole-mesh
        # solve handle of bench.distributed_solve_handle (row-range distributed V-cycle PCG over "peer memory", the
        # row partition chosen by solve_partition or forced to 'rcb' = renumbering + invariant aggregation), gather of
        # the local rows, partitioned adjoint; u against an undistributed solve, gradients against the oracle
        import bench
        world, size, min_dist = (int(a) for a in sys.argv[2:5])
        kind = sys.argv[5] if len(sys.argv) > 5 else 'auto'
        gmd = meshes.plate(size)
        owner = partition.rcb_owner(gmd.crds[:, :2], world)
        bar = threading.Barrier(world)
        shared = {'id': None, 'vals': [None] * world}
        out = [None] * world
        opts = nat.make_opts(rtol=1e-10, compliance=True, precond='multigrid', cheb_degree=1)

        def worker(rank):
            try:
                lm = partition.local_mesh(gmd, owner, rank, world)
                h = nat.Handle(lm.md.n_node, lm.md.cnct_quads, lm.md.cnct_beams, lm.md.known, device=0, n_row=lm.n_owned)

                def bcast(obj):
                    if rank == 0:
                        shared['id'] = obj
                    bar.wait()
                    got = shared['id']
                    bar.wait()
                    return got

                def allgather(obj):
                    shared.setdefault('box', [None] * world)[rank] = obj
                    bar.wait()
                    got = list(shared['box'])
                    bar.wait()
                    return got

                hd, smd, perm, info, levels = bench.distributed_solve_handle(nat, gmd, rank, world, 0, bcast, allgather,
                                                                             min_dist, use_p2p=True, partition_kind=kind)
                gu = D((smd.ndof,))
                fs = hd.forward(D.from_host(smd.crds), D.from_host(smd.prop_quads), D.from_host(smd.prop_beams),
                                D.from_host(smd.loads), gu, opts=opts)
                l2s = lm.l2g
                if perm is not None:
                    inv = np.empty(gmd.n_node, np.int64)
                    inv[perm] = np.arange(gmd.n_node)
                    l2s = inv[lm.l2g]
                uu, dc, dq = D((lm.md.ndof,)), D((lm.md.n_node, 3)), D((lm.md.n_quad, 5))
                nat.gather_rows(gu, D.from_host(l2s.astype(np.int32)), 6, out=uu)
                h.backward(D.from_host(lm.md.crds), D.from_host(lm.md.prop_quads), D.from_host(lm.md.prop_beams), uu, None,
                           dc, dq, None, opts=opts)
                ug = gu.download().reshape(-1, 6)
                if perm is not None:
                    tmp = np.empty_like(ug)
                    tmp[perm] = ug
                    ug = tmp
                ex, ar = hd.mg_dist_counters()
                out[rank] = (dict(info, pcg_iterations=fs.iterations, halo_exchanges=ex, scalar_allreduces=ar,
                                  peer_memory=bool(hd.mg_dist_p2p)),
                             lm.l2g[:lm.n_owned], dc.download()[:lm.n_owned], lm.quad_ids, dq.download(), ug.reshape(-1))
            except BaseException as e:      # the other ranks would wait for this one forever
                import traceback
                traceback.print_exc()
                print(f'EMU_RANK_ERROR rank {rank}: {e}', flush=True)
                os._exit(3)

        import jaxsso_b200.multigrid as mgmod
        _bh = mgmod.build_hierarchy
        mgmod.build_hierarchy = lambda rp, ci, max_coarse_nodes=64, **kw: _bh(rp, ci, max_coarse_nodes=8, **kw)  # small meshes
        _bi = dmg.build_hierarchy_invariant
        dmg.build_hierarchy_invariant = lambda rp, ci, perm, max_coarse_nodes=64, **kw: _bi(rp, ci, perm, max_coarse_nodes=8, **kw)
        th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        G, Q = np.zeros((gmd.n_node, 3)), np.zeros((gmd.n_quad, 5))
        for leg, ids, g, qi, q, ug in out:
            G[ids] = g
            Q[qi] = q
        us, its, _, _, _, _ = solve(gmd, 'multigrid', 1)
        rv, ru, rl, rdc, rdq, rdb = orc.value_and_grad(omesh(gmd))
        res = {'leg': out[0][0], 'g_err': float(np.abs(G - rdc).max() / np.abs(rdc).max()),
               'dq_err': float(np.abs(Q - rdq).max() / np.abs(rdq).max()), 'iters_single': its,
               'u_err_vs_single': float(max(np.linalg.norm(o[5] - us) for o in out) / np.linalg.norm(us)),
               'u_err_vs_oracle': float(np.linalg.norm(out[0][5] - ru) / np.linalg.norm(ru))}
    else:
        raise SystemExit('unknown mode')
    print('EMU_RESULT ' + json.dumps(res))


if __name__ == '__main__':
    main()
