"""Row-range distribution of the multigrid-preconditioned solve over N GPUs: host-side plans.

Why: on several GPUs the multigrid tier used to run replicated (every rank solved the whole system), so a
gradient evaluation did not get faster with N.  Here the *iteration* is distributed while the hierarchy stays
replicated:

  * the mesh is renumbered so that the nodes a rank owns are one contiguous range (`owner_permutation`,
    `renumber_mesh`); every rank holds a handle of the WHOLE renumbered mesh, assembles it and runs the numeric
    multigrid setup redundantly (no communication: 1.2 ms + 42 ms at 1M quads);
  * aggregates are numbered in ascending root order, so the coarse nodes "above" a rank's fine range are again
    (nearly) a contiguous range: `coarse_bounds` cuts every distributed level into N ranges;
  * in the V-cycle and the outer PCG a rank computes only ITS rows of every product (the existing kernels with
    offset pointers: `rowptr + s`, `y + 6 s`, full-length x) on full-length vectors in global numbering; before a
    product it receives the entries of x that its rows read but other ranks own (`level_halo`: the union of the
    columns of its rows of A_l, of its rows of P_l^T (restriction) and of its rows of P_{l-1} (prolongation));
  * levels with fewer than `min_dist_nodes` nodes run replicated: the restricted right-hand side is
    all-gathered once, everything below is redundant and the correction is known to every rank.

The plans are functions of the connectivity only (built once per model, identically on every rank, so the two
sides of each exchange agree without communication).  The device side is `jsso_mg_set_dist` +
`mg_solve_dist` (csrc/jsso_api.cu); `oracle/multigrid_ref.py::emulate_distributed_pcg` replays the same sequence
of range products and exchanges with NaN-poisoned ghosts on the CPU (tests/test_dist_multigrid.py).
"""
from __future__ import annotations

import numpy as np

from .meshes import MeshData


def owner_permutation(owner, n_rank):
    """new -> old node order with every rank's nodes contiguous (ascending old id inside a rank), and the
    range bounds (n_rank + 1)."""
    owner = np.asarray(owner)
    perm = np.argsort(owner, kind='stable').astype(np.int64)
    bounds = np.searchsorted(owner[perm], np.arange(n_rank + 1)).astype(np.int32)
    return perm, bounds


def renumber_mesh(md: MeshData, perm):
    """The same mesh with node `perm[k]` renamed k (elements keep their order)."""
    perm = np.asarray(perm, np.int64)
    inv = np.empty(md.n_node, np.int64)
    inv[perm] = np.arange(md.n_node)
    kn = md.known.astype(np.int64)
    return MeshData(crds=md.crds[perm], cnct_quads=inv[md.cnct_quads], prop_quads=md.prop_quads,
                    cnct_beams=inv[md.cnct_beams], prop_beams=md.prop_beams,
                    known=(6 * inv[kn // 6] + kn % 6).astype(np.int32),
                    loads=md.loads.reshape(-1, 6)[perm].reshape(-1), design_nodes=inv[md.design_nodes])


def coarse_bounds(lv, fine_bounds):
    """Range bounds of the coarse level of one coarsening step: an aggregate goes to the rank that owns its
    first (lowest) member, made monotone in the aggregate id (ids ascend with the root, so this moves only the
    few aggregates whose first member sits just across a range boundary, and the trailing singletons)."""
    first = lv['mem'][lv['mem_ptr'][:-1]]
    rk = np.searchsorted(fine_bounds, first, side='right') - 1
    rk = np.maximum.accumulate(rk)
    return np.searchsorted(rk, np.arange(fine_bounds.shape[0]), side='left').astype(np.int32)


def _cols_of_rows(rowptr, col, s, e):
    return col[rowptr[s]:rowptr[e]]


def level_halo(patterns, bounds):
    """need[r][s] = sorted node ids of this level that rank r reads and rank s owns (r != s).
    `patterns`: list of (rowptr, col, row_bounds) whose rows [row_bounds[r], row_bounds[r+1]) rank r computes."""
    n_rank = bounds.shape[0] - 1
    need = [[np.zeros(0, np.int32) for _ in range(n_rank)] for _ in range(n_rank)]
    for r in range(n_rank):
        cols = [_cols_of_rows(rp, ci, int(rb[r]), int(rb[r + 1])) for rp, ci, rb in patterns]
        ids = np.unique(np.concatenate(cols)) if cols else np.zeros(0, np.int32)
        ids = ids[(ids < bounds[r]) | (ids >= bounds[r + 1])]
        cut = np.searchsorted(ids, bounds)
        for s in range(n_rank):
            if s != r:
                need[r][s] = ids[cut[s]:cut[s + 1]].astype(np.int32)
    return need


def build_plan(rowptr, colidx, levels, fine_bounds, min_dist_nodes=20000, max_dist_levels=None):
    """Plan of the distributed solve.  Returns dict(n_dist, bounds[l] for l = 0..n_dist, need[l] for l < n_dist).
    Level l < n_dist is computed by row ranges; level n_dist is the first replicated one (it can be the dense
    coarsest level, index len(levels))."""
    fine_bounds = np.asarray(fine_bounds, np.int32)
    n_dist = 0
    for lv in levels:
        if lv['n_f'] >= min_dist_nodes and (max_dist_levels is None or n_dist < max_dist_levels):
            n_dist += 1
        else:
            break
    n_dist = max(n_dist, 1) if levels else 0     # the fine level is always distributed
    bounds = [fine_bounds]
    for l in range(n_dist):
        bounds.append(coarse_bounds(levels[l], bounds[l]))
    need = []
    for l in range(n_dist):
        a_rp, a_ci = (rowptr, colidx) if l == 0 else (levels[l - 1]['c_rowptr'], levels[l - 1]['c_col'])
        pats = [(a_rp, a_ci, bounds[l]),                                              # A_l x, rows of level l
                (levels[l]['pt_rowptr'], levels[l]['pt_col'], bounds[l + 1])]        # P_l^T r, rows of level l+1
        if l >= 1:
            pats.append((levels[l - 1]['p_rowptr'], levels[l - 1]['p_col'], bounds[l - 1]))   # P_{l-1} x_l
        need.append(level_halo(pats, bounds[l]))
    return dict(n_dist=n_dist, n_rank=fine_bounds.shape[0] - 1, bounds=bounds, need=need)


def rank_plan(plan, rank):
    """What `jsso_mg_set_dist` takes on one rank: per distributed level the peers and the packed
    send / receive node-id lists (a peer appears if either direction is non-empty)."""
    out = []
    for need in plan['need']:
        n_rank = len(need)
        peers = [s for s in range(n_rank) if s != rank and (need[rank][s].size or need[s][rank].size)]
        send = [need[s][rank] for s in peers]
        recv = [need[rank][s] for s in peers]
        cat = lambda xs: (np.concatenate(xs) if xs else np.zeros(0)).astype(np.int32)
        ptr = lambda xs: np.concatenate([[0], np.cumsum([x.size for x in xs])]).astype(np.int32)
        # peer-memory path: where my block starts in each peer's receive list (its peers in ascending rank order)
        remote_off = [int(sum(need[s][q].size for q in range(rank) if q != s)) for s in peers]
        out.append(dict(peer_rank=np.array(peers, np.int32), send_ptr=ptr(send), send_idx=cat(send),
                        recv_ptr=ptr(recv), recv_idx=cat(recv), remote_off=np.array(remote_off, np.int32)))
    return out


def common_max_recv(plan):
    """Largest number of nodes any rank receives in one exchange of any level (the stride of the peer-memory
    receive arenas, which must be the same on every rank)."""
    return int(max([sum(x.size for x in need[r]) for need in plan['need'] for r in range(plan['n_rank'])] + [1]))


def plan_summary(plan):
    """Sizes for logs / the bench line: rows per rank and exchanged nodes per level."""
    rows = [np.diff(b).tolist() for b in plan['bounds']]
    halo = [[int(sum(x.size for x in need[r])) for r in range(plan['n_rank'])] for need in plan['need']]
    return dict(n_dist=plan['n_dist'], rows_per_rank=rows, recv_nodes_per_rank=halo)
