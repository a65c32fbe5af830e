"""Distributed multigrid solve, host side (CPU): ownership renumbering, per-level row ranges and halo plans
(jaxsso_b200/dist_multigrid.py).  The sequence of range products and exchanges that `mg_solve_dist` runs on
the GPUs is replayed on simulated ranks with NaN-poisoned ghosts (oracle/multigrid_ref.py) and must
reproduce the global V-cycle / PCG; the packed per-rank lists are exchanged for real over `gloo`."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from jaxsso_b200 import dist_multigrid as dmg
from jaxsso_b200 import meshes, multigrid as mg, partition
from oracle import jaxsso_oracle as orc
from oracle import multigrid_ref as mgref
from tests.conftest import to_oracle_mesh
from tests.test_multigrid import scaled_system


def _setup(n, n_rank, min_dist_nodes, max_coarse_nodes=8):
    md0 = meshes.plate(n)
    owner = partition.rcb_owner(md0.crds[:, :2], n_rank)
    perm, bounds = dmg.owner_permutation(owner, n_rank)
    md = dmg.renumber_mesh(md0, perm)
    Ah, L, mask = scaled_system(md)
    rp, ci = Ah.indptr.astype(np.int32), Ah.indices.astype(np.int32)
    levels = mg.build_hierarchy(rp, ci, max_coarse_nodes=max_coarse_nodes)
    ref, Ac = mgref.reference_setup(Ah.data, Ah.indptr, Ah.indices, md.crds, mask, levels, Lt0=L.transpose(0, 2, 1))
    plan = dmg.build_plan(rp, ci, levels, bounds, min_dist_nodes=min_dist_nodes)
    return md0, md, perm, Ah, levels, ref, np.linalg.inv(Ac.toarray()), plan


def test_renumbering_is_a_symmetric_permutation_of_K():
    md0 = meshes.plate(8)
    owner = partition.rcb_owner(md0.crds[:, :2], 4)
    perm, bounds = dmg.owner_permutation(owner, 4)
    assert np.array_equal(np.diff(bounds), np.bincount(owner, minlength=4))
    for r in range(4):
        assert np.all(owner[perm[bounds[r]:bounds[r + 1]]] == r)
    md = dmg.renumber_mesh(md0, perm)
    K0 = orc.K_global(to_oracle_mesh(md0)).tocsr()
    K1 = orc.K_global(to_oracle_mesh(md)).tocsr()
    d = (6 * perm[:, None] + np.arange(6)[None, :]).ravel()
    assert abs(K1 - K0[d][:, d]).max() <= 1e-12 * abs(K0).max()
    assert np.array_equal(np.sort(d[md.known]), np.sort(md0.known))
    assert np.array_equal(md.loads, md0.loads[d])


@pytest.mark.parametrize('n_rank', [2, 4])
def test_coarse_ranges_and_plans_are_consistent(n_rank):
    _, md, _, Ah, levels, _, _, plan = _setup(24, n_rank, min_dist_nodes=40)
    assert plan['n_dist'] == 2 and len(plan['bounds']) == 3
    for l, b in enumerate(plan['bounds']):
        n_l = levels[l]['n_f'] if l < len(levels) else levels[-1]['n_c']
        assert b[0] == 0 and b[-1] == n_l and np.all(np.diff(b) >= 0)
    # most aggregates live where their members live
    lv = levels[0]
    own_f = np.searchsorted(plan['bounds'][0], np.arange(lv['n_f']), side='right') - 1
    own_c = np.searchsorted(plan['bounds'][1], lv['agg'], side='right') - 1
    assert np.mean(own_f == own_c) > 0.8
    rps = [dmg.rank_plan(plan, r) for r in range(n_rank)]
    for l in range(plan['n_dist']):
        for r in range(n_rank):
            me = rps[r][l]
            lo, hi = plan['bounds'][l][r], plan['bounds'][l][r + 1]
            assert np.all((me['send_idx'] >= lo) & (me['send_idx'] < hi))          # I only send what I own
            assert not np.any((me['recv_idx'] >= lo) & (me['recv_idx'] < hi))
            for i, p in enumerate(me['peer_rank']):
                other = rps[p][l]
                j = list(other['peer_rank']).index(r)
                sent = me['send_idx'][me['send_ptr'][i]:me['send_ptr'][i + 1]]
                recv = other['recv_idx'][other['recv_ptr'][j]:other['recv_ptr'][j + 1]]
                assert np.array_equal(sent, recv)
    s = dmg.plan_summary(plan)
    assert s['n_dist'] == 2 and sum(s['rows_per_rank'][0]) == md.n_node


@pytest.mark.parametrize('n_rank,min_dist,deg', [(2, 40, 1), (2, 40, 2), (4, 40, 1), (4, 200, 2), (4, 5, 1)])
def test_replayed_distributed_vcycle_equals_global(n_rank, min_dist, deg):
    _, md, _, Ah, levels, ref, Ainv, plan = _setup(24, n_rank, min_dist_nodes=min_dist)
    assert plan['n_dist'] == {40: 2, 200: 1, 5: len(levels)}[min_dist]
    b = np.random.default_rng(3).standard_normal(Ah.shape[0])
    z_ref = mgref.reference_vcycle(ref, Ainv, b, deg=deg)
    z, n_ex = mgref.emulate_distributed_vcycle(ref, Ainv, plan, b, deg=deg)
    assert not np.isnan(z).any()
    assert np.linalg.norm(z - z_ref) <= 1e-12 * np.linalg.norm(z_ref)
    # exchanges per V-cycle: per distributed level 2 (residual, restriction) + deg (post-smoother) + (deg - 1)
    # (pre-smoother) + 1 for the correction of every distributed coarse level
    nd = plan['n_dist']
    assert n_ex == nd * (2 + deg + deg - 1) + (nd - 1)


def test_a_missing_ghost_is_detected():
    """The NaN poisoning works: dropping one received node from the plan breaks the replay."""
    _, md, _, Ah, levels, ref, Ainv, plan = _setup(24, 2, min_dist_nodes=200)
    b = np.random.default_rng(3).standard_normal(Ah.shape[0])
    plan['need'][0][1][0] = plan['need'][0][1][0][1:]
    z, _ = mgref.emulate_distributed_vcycle(ref, Ainv, plan, b, deg=1)
    assert np.isnan(z).any()


@pytest.mark.parametrize('n_rank', [2, 4])
def test_replayed_distributed_pcg_solves_the_original_system(n_rank):
    md0, md, perm, Ah, levels, ref, Ainv, plan = _setup(24, n_rank, min_dist_nodes=40)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(Ah.shape[0])
    # global PCG with the plain V-cycle
    x = np.zeros_like(b); r = b.copy()
    z = mgref.reference_vcycle(ref, Ainv, r, deg=1)
    p = z.copy(); rz = r @ z
    for it in range(1, 200):
        q = Ah @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        if np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b):
            break
        z = mgref.reference_vcycle(ref, Ainv, r, deg=1)
        rz, rz_old = r @ z, rz
        p = z + (rz / rz_old) * p
    xd, itd = mgref.emulate_distributed_pcg(ref, Ainv, plan, b, deg=1, rtol=1e-8)
    assert itd == it and itd < 100
    assert np.linalg.norm(xd - x) <= 1e-9 * np.linalg.norm(x)


def test_block_ordering_keeps_the_iteration_count():
    """Renumbering by owner changes the greedy aggregates; the preconditioner stays as good."""
    def iters(md):
        Ah, L, mask = scaled_system(md)
        levels = mg.build_hierarchy(Ah.indptr.astype(np.int32), Ah.indices.astype(np.int32), max_coarse_nodes=30)
        ref, Ac = mgref.reference_setup(Ah.data, Ah.indptr, Ah.indices, md.crds, mask, levels, Lt0=L.transpose(0, 2, 1))
        Ainv = np.linalg.inv(Ac.toarray())
        b = np.random.default_rng(1).standard_normal(Ah.shape[0])
        x = np.zeros_like(b); r = b.copy()
        z = mgref.reference_vcycle(ref, Ainv, r, deg=1)
        p = z.copy(); rz = r @ z
        for it in range(1, 300):
            q = Ah @ p
            a = rz / (p @ q)
            x += a * p; r -= a * q
            if np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b):
                return it
            z = mgref.reference_vcycle(ref, Ainv, r, deg=1)
            rz, rz_old = r @ z, rz
            p = z + (rz / rz_old) * p
        return 300
    md0 = meshes.plate(32)
    base = iters(md0)
    for n_rank in (2, 8):
        perm, _ = dmg.owner_permutation(partition.rcb_owner(md0.crds[:, :2], n_rank), n_rank)
        assert iters(dmg.renumber_mesh(md0, perm)) <= base * 1.25 + 2


def _pcg_iterations(Ah, ref, Ainv, seed=1):
    b = np.random.default_rng(seed).standard_normal(Ah.shape[0])
    x = np.zeros_like(b); r = b.copy()
    z = mgref.reference_vcycle(ref, Ainv, r, deg=1)
    p = z.copy(); rz = r @ z
    for it in range(1, 300):
        q = Ah @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        if np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b):
            return it, x
        z = mgref.reference_vcycle(ref, Ainv, r, deg=1)
        rz, rz_old = r @ z, rz
        p = z + (rz / rz_old) * p
    return 300, x


@pytest.mark.parametrize('n_rank', [2, 8])
def test_invariant_hierarchy_has_the_aggregates_of_the_original_numbering(n_rank):
    """`build_hierarchy_invariant`: on the owner-renumbered mesh the aggregates of EVERY level are those the
    single-GPU hierarchy forms on the original numbering (as sets of original nodes), coarse ranges stay contiguous
    per rank, and the PCG takes the same number of iterations as on the original mesh (the plain renumbered
    hierarchy does not)."""
    md0 = meshes.plate(32)
    A0, L0, mask0 = scaled_system(md0)
    lev0 = mg.build_hierarchy(A0.indptr.astype(np.int32), A0.indices.astype(np.int32), max_coarse_nodes=30)
    perm, bounds = dmg.owner_permutation(partition.rcb_owner(md0.crds[:, :2], n_rank), n_rank)
    md = dmg.renumber_mesh(md0, perm)
    Ah, L, mask = scaled_system(md)
    rp, ci = Ah.indptr.astype(np.int32), Ah.indices.astype(np.int32)
    lev = dmg.build_hierarchy_invariant(rp, ci, perm, max_coarse_nodes=30)
    assert [l['n_c'] for l in lev] == [l['n_c'] for l in lev0]
    # level 0: same partition of the original nodes into aggregates
    sets0 = {frozenset(np.flatnonzero(lev0[0]['agg'] == a).tolist()) for a in range(lev0[0]['n_c'])}
    sets1 = {frozenset(perm[np.flatnonzero(lev[0]['agg'] == a)].tolist()) for a in range(lev[0]['n_c'])}
    assert sets0 == sets1
    # coarse ranges per rank are contiguous and cover the level; first members ascend with the coarse id
    b = bounds
    for l in lev:
        first = l['mem'][l['mem_ptr'][:-1]]
        assert np.all(np.diff(first) > 0)
        b = dmg.coarse_bounds(l, b)
        assert b[0] == 0 and b[-1] == l['n_c'] and np.all(np.diff(b) >= 0)
    ref0, Ac0 = mgref.reference_setup(A0.data, A0.indptr, A0.indices, md0.crds, mask0, lev0, Lt0=L0.transpose(0, 2, 1))
    ref1, Ac1 = mgref.reference_setup(Ah.data, Ah.indptr, Ah.indices, md.crds, mask, lev, Lt0=L.transpose(0, 2, 1))
    it0, x0 = _pcg_iterations(A0, ref0, np.linalg.inv(Ac0.toarray()))
    # the same right-hand side in the renumbered dof order
    d = (6 * perm[:, None] + np.arange(6)[None, :]).ravel()
    b0 = np.random.default_rng(1).standard_normal(A0.shape[0])
    Ainv1 = np.linalg.inv(Ac1.toarray())
    x = np.zeros(Ah.shape[0]); r = b0[d].copy()
    z = mgref.reference_vcycle(ref1, Ainv1, r, deg=1)
    p = z.copy(); rz = r @ z
    for it1 in range(1, 300):
        q = Ah @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        if np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b0):
            break
        z = mgref.reference_vcycle(ref1, Ainv1, r, deg=1)
        rz, rz_old = r @ z, rz
        p = z + (rz / rz_old) * p
    assert abs(it1 - it0) <= 1
    assert np.linalg.norm(x - x0[d]) <= 1e-7 * np.linalg.norm(x0)


def test_structured_plate_is_partitioned_without_renumbering():
    """`solve_partition`: a grid numbered row by row is cut into contiguous ranges of its own numbering (identity
    permutation: the distributed solve then has the single-GPU hierarchy); a scrambled numbering falls back to RCB."""
    md = meshes.plate(32)
    Ah, _, _ = scaled_system(meshes.plate(8))      # only the pattern matters; use a small one for the scrambled case
    from jaxsso_b200 import _native as nat
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=nat.JSSO_DEVICE_NONE)
    rp, ci = h.pattern()
    perm, bounds, kind = dmg.solve_partition(md.crds, rp, ci, 4, max_halo_fraction=0.3)
    assert kind == 'natural' and np.array_equal(perm, np.arange(md.n_node))
    assert np.array_equal(bounds, dmg.natural_bounds(md.n_node, 4)) and dmg.halo_fraction(rp, ci, bounds) < 0.3
    rng = np.random.default_rng(0)
    scr = rng.permutation(md.n_node)
    md2 = dmg.renumber_mesh(md, scr)
    h2 = nat.Handle(md2.n_node, md2.cnct_quads, md2.cnct_beams, md2.known, device=nat.JSSO_DEVICE_NONE)
    rp2, ci2 = h2.pattern()
    perm2, bounds2, kind2 = dmg.solve_partition(md2.crds, rp2, ci2, 4, max_halo_fraction=0.3)
    assert kind2 == 'rcb' and bounds2[-1] == md.n_node


# ------------------------------------------------------------------------------- gloo, world_size 2
def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    md0, md, perm, Ah, levels, ref, Ainv, plan = _setup(16, world, min_dist_nodes=30)
    mine = dmg.rank_plan(plan, rank)
    ok = True
    for l in range(plan['n_dist']):
        A = ref[l].A
        n = A.shape[0]
        xg = np.random.default_rng(10 + l).standard_normal(n)
        lo, hi = 6 * plan['bounds'][l][rank], 6 * plan['bounds'][l][rank + 1]
        x = np.full(n, np.nan)
        x[lo:hi] = xg[lo:hi]
        me = mine[l]
        reqs, recvs = [], []
        for i, p in enumerate(me['peer_rank']):
            s_ids = me['send_idx'][me['send_ptr'][i]:me['send_ptr'][i + 1]]
            r_ids = me['recv_idx'][me['recv_ptr'][i]:me['recv_ptr'][i + 1]]
            if s_ids.size:     # pack -> send
                reqs.append(dist.isend(torch.from_numpy(x.reshape(-1, 6)[s_ids].copy()), int(p)))
            if r_ids.size:
                rb = torch.zeros((r_ids.size, 6), dtype=torch.float64)
                reqs.append(dist.irecv(rb, int(p)))
                recvs.append((r_ids, rb))
        for rq in reqs:
            rq.wait()
        for r_ids, rb in recvs:   # unpack
            x.reshape(-1, 6)[r_ids] = rb.numpy()
        y = A[lo:hi] @ x
        ok = ok and bool(np.allclose(y, (A @ xg)[lo:hi], rtol=1e-13, atol=1e-13 * abs(A @ xg).max()))
    ret[rank] = ok
    dist.destroy_process_group()


def test_rank_plans_over_gloo():
    import torch.multiprocessing as tmp
    world = 2
    ctx = tmp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs)
    assert all(ret[r] for r in range(world))


@pytest.mark.parametrize('n_rank', [2, 4, 8])
def test_ghost_rows_of_the_corrected_iterate_read_only_delivered_data(n_rank):
    """The fused distributed V-cycle recomputes x1 = b / theta + P x_c on the ghost rows of every distributed level
    instead of exchanging it (`plan['ghost']`).  Pure bookkeeping check of the plan: (i) the ghost rows of a rank are
    exactly the foreign columns its own rows of A_l read -- what the post-smoother gathers; (ii) every one of them is in
    the halo of the level's exchanges, so b has arrived there; (iii) every coarse column of a ghost row's prolongator is
    either owned by the rank or in the halo of level l + 1 (distributed) -- x_c is there -- or level l + 1 is replicated."""
    _, md, _, Ah, levels, _, _, plan = _setup(24, n_rank, min_dist_nodes=40)
    rp, ci = Ah.indptr.astype(np.int32), Ah.indices.astype(np.int32)
    n_dist = plan['n_dist']
    assert n_dist == 2
    for l in range(n_dist):
        a_rp, a_ci = (rp, ci) if l == 0 else (levels[l - 1]['c_rowptr'], levels[l - 1]['c_col'])
        lv = levels[l]
        for r in range(n_rank):
            lo, hi = int(plan['bounds'][l][r]), int(plan['bounds'][l][r + 1])
            cols = np.unique(a_ci[a_rp[lo]:a_rp[hi]]) if hi > lo else np.zeros(0, np.int32)
            foreign = cols[(cols < lo) | (cols >= hi)]
            ghost = np.asarray(plan['ghost'][l][r])
            assert np.array_equal(np.sort(ghost), foreign)                                      # (i)
            halo_l = np.concatenate([np.asarray(plan['need'][l][r][s]) for s in range(n_rank)] + [np.zeros(0, np.int32)])
            assert np.all(np.isin(ghost, halo_l))                                               # (ii)
            if ghost.size and l + 1 < n_dist:                                                   # (iii)
                pc = np.unique(np.concatenate([lv['p_col'][lv['p_rowptr'][g]:lv['p_rowptr'][g + 1]] for g in ghost]))
                lo1, hi1 = int(plan['bounds'][l + 1][r]), int(plan['bounds'][l + 1][r + 1])
                halo_1 = np.concatenate([np.asarray(plan['need'][l + 1][r][s]) for s in range(n_rank)] + [np.zeros(0, np.int32)])
                missing = pc[((pc < lo1) | (pc >= hi1)) & ~np.isin(pc, halo_1)]
                assert missing.size == 0
        rps = [dmg.rank_plan(plan, r) for r in range(n_rank)]
        for r in range(n_rank):
            assert np.array_equal(np.sort(rps[r][l]['ghost_rows']), np.sort(np.asarray(plan['ghost'][l][r])))
