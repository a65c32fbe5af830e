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


def natural_bounds(n_node, n_rank):
    """Equal contiguous ranges of the mesh's own numbering (no renumbering at all)."""
    return np.array([(n_node * r) // n_rank for r in range(n_rank + 1)], np.int32)


def halo_fraction(rowptr, colidx, bounds):
    """Largest ratio (nodes a rank reads from other ranks) / (nodes it owns) over the ranks, for row ranges
    `bounds` of the block-CSR pattern: small for a banded numbering cut into contiguous ranges."""
    worst = 0.0
    for r in range(bounds.shape[0] - 1):
        s, e = int(bounds[r]), int(bounds[r + 1])
        if e <= s:
            continue
        cols = np.unique(colidx[rowptr[s]:rowptr[e]])
        worst = max(worst, float(np.count_nonzero((cols < s) | (cols >= e))) / (e - s))
    return worst


def solve_partition(crds, rowptr, colidx, n_rank, max_halo_fraction=0.05):
    """Row partition of the distributed multigrid solve: (perm, bounds, kind).

    * 'natural': the mesh numbering is banded enough (structured grids numbered row by row, meshes already ordered
      by a bandwidth-reducing permutation) that equal contiguous ranges of it have small halos -- no renumbering, so
      the aggregates, the hierarchy and hence the PCG iteration count are EXACTLY those of the single-GPU solve;
    * 'rcb': recursive coordinate bisection + ownership renumbering (`owner_permutation`); use
      `build_hierarchy_invariant` so that the aggregates are still formed in the mesh's own order."""
    from . import partition
    n = rowptr.shape[0] - 1
    nb = natural_bounds(n, n_rank)
    if halo_fraction(rowptr, colidx, nb) <= max_halo_fraction:
        return np.arange(n, dtype=np.int64), nb, 'natural'
    owner = partition.rcb_owner(np.asarray(crds)[:, :2], n_rank)
    perm, bounds = owner_permutation(owner, n_rank)
    return perm, bounds, 'rcb'


def _permute_pattern(rowptr, colidx, new_of_old):
    """Block-CSR pattern with node `o` renamed `new_of_old[o]` (columns sorted per row)."""
    import scipy.sparse as sp
    n = rowptr.shape[0] - 1
    A = sp.csr_matrix((np.ones(colidx.shape[0], np.int8), colidx, rowptr), shape=(n, n)).tocoo()
    B = sp.csr_matrix((A.data, (new_of_old[A.row], new_of_old[A.col])), shape=(n, n))
    B.sort_indices()
    return B.indptr.astype(np.int32), B.indices.astype(np.int32)


def build_hierarchy_invariant(rowptr_perm, colidx_perm, perm, max_coarse_nodes=64, max_levels=12):
    """Smoothed-aggregation hierarchy for a RENUMBERED mesh (`perm`: new -> old node order, e.g. from
    `owner_permutation`) whose aggregates are those of the ORIGINAL numbering: the greedy aggregation
    (`multigrid.aggregate`) is order dependent, so aggregating the renumbered graph gives a different -- in
    practice worse -- hierarchy for every rank count (1M quads: 164 PCG iterations on one GPU, 234 on 8 ranks).
    Here every level is aggregated in its original order and only then renumbered: coarse nodes are numbered by
    their first member in the renumbered fine order, which keeps every rank's coarse nodes contiguous
    (`coarse_bounds`).  The hierarchy is the single-GPU one up to numbering; iteration counts agree to rounding."""
    from . import multigrid
    perm = np.asarray(perm, np.int64)
    n = perm.shape[0]
    new_of_old = np.empty(n, np.int64)
    new_of_old[perm] = np.arange(n)
    rp_p, ci_p = np.asarray(rowptr_perm, np.int32), np.asarray(colidx_perm, np.int32)
    old_of_new = perm                                        # level-l node: new id -> original id
    levels = []
    while rp_p.shape[0] - 1 > max_coarse_nodes and len(levels) < max_levels:
        n_f = rp_p.shape[0] - 1
        inv = np.empty(n_f, np.int64)
        inv[old_of_new] = np.arange(n_f)                     # original id -> new id
        rp_o, ci_o = _permute_pattern(rp_p, ci_p, old_of_new)   # the level's graph in its original numbering
        agg_o, n_c = multigrid.aggregate(rp_o, ci_o)         # aggregates as the single-GPU hierarchy forms them
        if n_c >= n_f:
            break
        agg_by_new = agg_o[old_of_new]                       # original coarse id of every (new-numbered) fine node
        first = np.full(n_c, n_f, np.int64)
        np.minimum.at(first, agg_by_new, np.arange(n_f))
        c_old_of_new = np.argsort(first, kind='stable')      # coarse: new id -> original id
        c_new_of_old = np.empty(n_c, np.int64)
        c_new_of_old[c_old_of_new] = np.arange(n_c)
        lv = multigrid.build_level(rp_p, ci_p, c_new_of_old[agg_by_new].astype(np.int32), n_c)
        levels.append(lv)
        rp_p, ci_p, old_of_new = lv['c_rowptr'], lv['c_col'], c_old_of_new
    return levels


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


def level_halo(patterns, bounds, extra=None):
    """need[r][s] = sorted node ids of this level that rank r reads and rank s owns (r != s).
    `patterns`: list of (rowptr, col, row_bounds) whose rows [row_bounds[r], row_bounds[r+1]) rank r computes;
    `extra[r]`: further ids rank r reads (columns of rows it recomputes as ghost rows)."""
    n_rank = bounds.shape[0] - 1
    need = [[np.zeros(0, np.int32) for _ in range(n_rank)] for _ in range(n_rank)]
    for r in range(n_rank):
        cols = [_cols_of_rows(rp, ci, int(rb[r]), int(rb[r + 1])) for rp, ci, rb in patterns]
        if extra is not None:
            cols.append(np.asarray(extra[r]))
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
    n_rank_ = fine_bounds.shape[0] - 1
    # ghost rows of level l: the rows other ranks own whose entries this rank's rows of A_l read.  The corrected
    # iterate x1 = b / theta + P x_c of the fused V-cycle is RECOMPUTED on them (b is there after its exchange, the
    # prolongator rows are local data) instead of being exchanged: one synchronisation point fewer per level and
    # V-cycle, for a slightly larger exchange of x_c (the coarse columns of the ghost rows' prolongator)
    ghost = []
    for l in range(n_dist):
        a_rp, a_ci = (rowptr, colidx) if l == 0 else (levels[l - 1]['c_rowptr'], levels[l - 1]['c_col'])
        gl = []
        for r in range(n_rank_):
            s_, e_ = int(bounds[l][r]), int(bounds[l][r + 1])
            ids = np.unique(a_ci[a_rp[s_]:a_rp[e_]]) if e_ > s_ else np.zeros(0, np.int32)
            gl.append(ids[(ids < s_) | (ids >= e_)].astype(np.int32))
        ghost.append(gl)
    need = []
    for l in range(n_dist):
        a_rp, a_ci = (rowptr, colidx) if l == 0 else (levels[l - 1]['c_rowptr'], levels[l - 1]['c_col'])
        pats = [(a_rp, a_ci, bounds[l]),                                              # A_l x, rows of level l
                (levels[l]['pt_rowptr'], levels[l]['pt_col'], bounds[l + 1])]        # P_l^T r, rows of level l+1
        extra = None
        if l >= 1:
            pats.append((levels[l - 1]['p_rowptr'], levels[l - 1]['p_col'], bounds[l - 1]))   # P_{l-1} x_l
            pl = levels[l - 1]                                                        # ... and on the ghost rows of level l-1
            extra = [pl['p_col'][_concat_ranges(pl['p_rowptr'], ghost[l - 1][r])] for r in range(n_rank_)]
        need.append(level_halo(pats, bounds[l], extra))
    # the first replicated level: its restricted right-hand side is all-gathered once per V-cycle; as an exchange
    # plan "every rank needs every other rank's whole range" it runs over the same peer-memory push / wait kernels
    n_rank = fine_bounds.shape[0] - 1
    gb = bounds[n_dist]
    gather = [[(np.arange(gb[s], gb[s + 1], dtype=np.int32) if s != r else np.zeros(0, np.int32)) for s in range(n_rank)]
              for r in range(n_rank)]
    return dict(n_dist=n_dist, n_rank=n_rank, bounds=bounds, need=need, gather=gather, ghost=ghost)


def rank_plan(plan, rank):
    """What `jsso_mg_set_dist` takes on one rank: per distributed level the peers and the packed
    send / receive node-id lists (a peer appears if either direction is non-empty)."""
    out = []
    for need in list(plan['need']) + ([plan['gather']] if plan.get('gather') is not None else []):
        n_rank = len(need)
        peers = [s for s in range(n_rank) if s != rank and (need[rank][s].size or need[s][rank].size)]
        send = [need[s][rank] for s in peers]
        recv = [need[rank][s] for s in peers]
        cat = lambda xs: (np.concatenate(xs) if xs else np.zeros(0)).astype(np.int32)
        ptr = lambda xs: np.concatenate([[0], np.cumsum([x.size for x in xs])]).astype(np.int32)
        # peer-memory path: where my block starts in each peer's receive list (its peers in ascending rank order)
        remote_off = [int(sum(need[s][q].size for q in range(rank) if q != s)) for s in peers]
        out.append(dict(peer_rank=np.array(peers, np.int32), send_ptr=ptr(send), send_idx=cat(send),
                        recv_ptr=ptr(recv), recv_idx=cat(recv), remote_off=np.array(remote_off, np.int32),
                        ghost_rows=np.zeros(0, np.int32)))
    for l, gl in enumerate(plan.get('ghost') or []):
        out[l]['ghost_rows'] = np.ascontiguousarray(gl[rank], np.int32)
    return out


def _concat_ranges(ptr, ids):
    """Concatenation of arange(ptr[i], ptr[i + 1]) for i in ids (vectorised)."""
    ids = np.asarray(ids, np.int64)
    cnt = (ptr[ids + 1] - ptr[ids]).astype(np.int64)
    tot = int(cnt.sum())
    if tot == 0:
        return np.zeros(0, np.int64)
    start = np.repeat(ptr[ids].astype(np.int64), cnt)
    return start + (np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt))


def _slot_hull(ptr, s, e, extra_rows):
    lo, hi = int(ptr[s]), int(ptr[e])
    if len(extra_rows):
        lo = min(lo, int(ptr[int(np.min(extra_rows))]))
        hi = max(hi, int(ptr[int(np.max(extra_rows)) + 1]))
    return lo, hi


def setup_plan(rowptr, colidx, levels, plan, rank, drop_ghost=False):
    """What ONE rank has to compute of the numeric multigrid setup at the distributed levels (jsso_mg_set_dist_setup).

    Level l < n_dist, own fine rows [fs, fe), own coarse rows [cs, ce):
      * Galerkin blocks Ac of the own coarse rows (a contiguous slot range; all-gathered afterwards, so that the next
        level sees its whole matrix);
      * the AP blocks those Ac blocks read (`ap_slots`);
      * the P blocks read by those AP blocks, by those Ac blocks, by the restriction rows P^T[cs:ce] and by the
        prolongation rows P[fs:fe] (`p_slots`) -- rows of neighbouring ranks are recomputed (ghost rows), which
        needs no communication because every rank can assemble any row of the fine matrix.
    Level 0 also gets the row hulls the fine matrix must be assembled / scaled on: `scale_rows` = every row whose
    scaled blocks are read (own rows, rows of p_slots and ap_slots), `factor_rows` = those plus their columns (the
    block-Jacobi factors W_r, W_c of a scaled block)."""
    n_dist = plan['n_dist']
    out = []
    for l in range(n_dist):
        lv = levels[l]
        fs, fe = int(plan['bounds'][l][rank]), int(plan['bounds'][l][rank + 1])
        cs, ce = int(plan['bounds'][l + 1][rank]), int(plan['bounds'][l + 1][rank + 1])
        ac0, ac1 = int(lv['c_rowptr'][cs]), int(lv['c_rowptr'][ce])
        terms = np.arange(lv['cl_ptr'][ac0], lv['cl_ptr'][ac1], dtype=np.int64)
        ap_slots = np.unique(lv['cl_ap'][terms])
        ap_terms = _concat_ranges(lv['apl_ptr'], ap_slots)
        p_slots = np.unique(np.concatenate([
            lv['cl_p'][terms].astype(np.int64), lv['apl_p'][ap_terms].astype(np.int64),
            lv['pt_src'][lv['pt_rowptr'][cs]:lv['pt_rowptr'][ce]].astype(np.int64),
            np.arange(lv['p_rowptr'][fs], lv['p_rowptr'][fe], dtype=np.int64),
            _concat_ranges(lv['p_rowptr'], (plan.get('ghost') or [[np.zeros(0, np.int32)] * plan['n_rank']] * n_dist)[l][rank])]))
        if drop_ghost and l == 0 and rank == 0:     # test hook: a ghost block missing from the plan must be noticed
            ghost = p_slots[p_slots >= lv['p_rowptr'][fe]]
            p_slots = np.setdiff1d(p_slots, ghost[-1:])
        ac_bounds = lv['c_rowptr'][plan['bounds'][l + 1]].astype(np.int32)
        d = dict(p_slots=p_slots.astype(np.int32), ap_slots=ap_slots.astype(np.int32), ac_bounds=ac_bounds,
                 # slot hull of the prolongation rows this rank applies (own rows + ghost rows): the FP32 copy
                 p_own=_slot_hull(lv['p_rowptr'], fs, fe, (plan.get('ghost') or [[np.zeros(0, np.int32)] * plan['n_rank']] * n_dist)[l][rank]),
                 pt_own=(int(lv['pt_rowptr'][cs]), int(lv['pt_rowptr'][ce])))
        if l == 0:
            n_f = lv['n_f']
            p_row = np.repeat(np.arange(n_f, dtype=np.int64), np.diff(lv['p_rowptr']))
            ap_row = np.repeat(np.arange(n_f, dtype=np.int64), np.diff(lv['ap_rowptr']))
            rows = np.concatenate([[fs, max(fe - 1, fs)], p_row[p_slots], ap_row[ap_slots]])
            lo, hi = int(rows.min()), int(rows.max()) + 1
            cols = colidx[rowptr[lo]:rowptr[hi]]
            wlo, whi = min(lo, int(cols.min())), max(hi, int(cols.max()) + 1)
            d['scale_rows'] = (lo, hi)
            d['factor_rows'] = (wlo, whi)
        out.append(d)
    return out


def common_max_recv(plan):
    """Largest number of nodes any rank receives in one exchange of any level (the stride of the peer-memory
    receive arenas, which must be the same on every rank)."""
    needs = list(plan['need']) + ([plan['gather']] if plan.get('gather') is not None else [])
    return int(max([sum(x.size for x in need[r]) for need in needs for r in range(plan['n_rank'])] + [1]))


def plan_summary(plan):
    """Sizes for logs / the bench line: rows per rank and exchanged nodes per level."""
    rows = [np.diff(b).tolist() for b in plan['bounds']]
    halo = [[int(sum(x.size for x in need[r])) for r in range(plan['n_rank'])] for need in plan['need']]
    return dict(n_dist=plan['n_dist'], rows_per_rank=rows, recv_nodes_per_rank=halo)
