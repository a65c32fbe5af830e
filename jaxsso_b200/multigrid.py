"""Smoothed-aggregation multigrid hierarchy for the block-CSR stiffness matrix: the symbolic
(connectivity-only) part, computed once per model on the host.  The numeric part runs on the GPU
(csrc/jsso_multigrid.cuh); its SciPy restatement used by the tests lives in oracle/multigrid_ref.py.

Why: block-Jacobi CG needs 6.5e4 iterations at 1M quads (measured 64 765 on
a B200); with a V-cycle of smoothed aggregation on rigid-body modes the count is 40-100 and
nearly mesh independent.  This is not part of the reference (it solves with
SuperLU, JaxSSO/solver.py:195-197); it only changes how fast the same u is reached.

Per level l (fine matrix A_l in 6x6 block-CSR over n_l nodes with coordinates X_l):
  * aggregates: greedy root + neighbours on the node graph, leftovers join a neighbour;
  * tentative prolongator: node i maps to its aggregate a with the rigid-body block
        T_i = [[I, -[r]x], [0, I]],  r = X_i - centroid_a          (rows of prescribed dofs zeroed)
    (a coarse node carries a translation and a rotation about its centroid);
  * smoothed prolongator P = (I - omega D^-1 A) T, omega = 4 / (3 lambda_max(D^-1 A));
  * Galerkin operator A_{l+1} = P^T A P, computed as AP = A P, then P^T (AP).
Every sparse product is a *gather*: each output block owns a list of (left slot, right slot)
pairs fixed by the connectivity, so the numeric kernels need no atomics and sum in a fixed order.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- symbolic
def aggregate(rowptr, colidx):
    """Greedy aggregation (Vanek et al.) by the native host routine `jsso_mg_aggregate`: a node whose
    whole neighbourhood is free roots an aggregate of itself + neighbours; leftovers join the aggregate of
    their first aggregated neighbour (or become singletons).  Deterministic (ascending node order); the
    test suite checks it against a plain Python restatement (oracle/multigrid_ref.py)."""
    import ctypes
    from . import _native as nat
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colidx = np.ascontiguousarray(colidx, np.int32)
    n = rowptr.shape[0] - 1
    agg = np.empty(n, np.int32)
    na = ctypes.c_int32()
    rc = nat.lib().jsso_mg_aggregate(n, nat._ptr(rowptr), nat._ptr(colidx), nat._ptr(agg), ctypes.byref(na))
    if rc:
        raise nat.JssoError(rc, 'jsso_mg_aggregate')
    return agg, int(na.value)


def _csr_lists(keys_major, payload_cols, n_major):
    """Group payload rows by `keys_major` (already sorted) -> (ptr, payload)."""
    ptr = np.zeros(n_major + 1, np.int64)
    np.add.at(ptr, keys_major + 1, 1)
    return np.cumsum(ptr).astype(np.int32), payload_cols


def _pattern_and_lists_native(row, col, left, right, n_row):
    """The same through the native host routine (stable two-level sort; input already in (left, right) order)."""
    import ctypes
    from . import _native as nat
    arrs = [np.ascontiguousarray(a, np.int32) for a in (row, col, left, right)]
    m = arrs[0].shape[0]
    rowptr = np.empty(n_row + 1, np.int32)
    ocol, ptr = np.empty(max(m, 1), np.int32), np.empty(m + 1, np.int32)
    lo, ro = np.empty(max(m, 1), np.int32), np.empty(max(m, 1), np.int32)
    nb = ctypes.c_int64()
    rc = nat.lib().jsso_mg_pattern_lists(m, *(nat._ptr(a) for a in arrs), n_row, nat._ptr(rowptr), nat._ptr(ocol),
                                         nat._ptr(ptr), nat._ptr(lo), nat._ptr(ro), ctypes.byref(nb))
    if rc:
        raise nat.JssoError(rc, 'jsso_mg_pattern_lists')
    nb = int(nb.value)
    return rowptr, ocol[:nb].copy(), ptr[:nb + 1].copy(), lo[:m], ro[:m]


def _pattern_and_lists(row, col, left, right, n_row, n_col):
    """Triples (row, col, left slot, right slot) -> sorted block pattern (rowptr, colidx) and, per
    output block, its (left, right) list in a deterministic order."""
    # every caller emits its triples in (left, right) order already, so a STABLE sort by (row, col) alone is
    # the lexicographic (row, col, left, right) order: native routine; the three-key sort is kept for other input
    dl = np.diff(left)
    if np.all((dl > 0) | ((dl == 0) & (np.diff(right) >= 0))):
        return _pattern_and_lists_native(row, col, left, right, n_row)
    key = row.astype(np.int64) * n_col + col
    order = np.lexsort((right, left, key))
    key, left, right = key[order], left[order], right[order]
    uniq, start = np.unique(key, return_index=True)
    ptr = np.append(start, key.shape[0]).astype(np.int32)
    orow = (uniq // n_col).astype(np.int32)
    ocol = (uniq % n_col).astype(np.int32)
    rowptr = np.zeros(n_row + 1, np.int64)
    np.add.at(rowptr, orow + 1, 1)
    return (np.cumsum(rowptr).astype(np.int32), ocol, ptr, left.astype(np.int32), right.astype(np.int32))


def build_level(rowptr, colidx, agg, n_c):
    """All connectivity-only data of one coarsening step (int32 arrays)."""
    n_f = rowptr.shape[0] - 1
    nnz = colidx.shape[0]
    brow = np.repeat(np.arange(n_f, dtype=np.int32), np.diff(rowptr))
    slot = np.arange(nnz, dtype=np.int32)
    # P pattern + smoothing lists: P[i, agg(j)] gets A[i,j] T_j for every stored (i,j)
    p_rowptr, p_col, ps_ptr, ps_a, ps_j = _pattern_and_lists(brow, agg[colidx], slot, colidx, n_f, n_c)
    nnz_p = p_col.shape[0]
    prow = np.repeat(np.arange(n_f, dtype=np.int32), np.diff(p_rowptr))
    p_own = (p_col == agg[prow]).astype(np.int32)          # tentative block present
    # AP[i,d] = sum_j A[i,j] P[j,d]
    cnt = np.diff(p_rowptr)[colidx]                         # |cols(P_j)| for each A block (i,j)
    a_rep = np.repeat(slot, cnt)
    i_rep = np.repeat(brow, cnt)
    off = np.repeat(p_rowptr[colidx], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
    off = off.astype(np.int32)
    ap_rowptr, ap_col, apl_ptr, apl_a, apl_p = _pattern_and_lists(i_rep, p_col[off], a_rep, off, n_f, n_c)
    nnz_ap = ap_col.shape[0]
    # Ac[c,d] = sum_i P[i,c]^T AP[i,d]
    pslot = np.arange(nnz_p, dtype=np.int32)
    cnt2 = np.diff(ap_rowptr)[prow]
    p_rep = np.repeat(pslot, cnt2)
    c_rep = np.repeat(p_col, cnt2)
    off2 = (np.repeat(ap_rowptr[prow], cnt2) +
            (np.arange(cnt2.sum()) - np.repeat(np.cumsum(cnt2) - cnt2, cnt2))).astype(np.int32)
    c_rowptr, c_col, cl_ptr, cl_p, cl_ap = _pattern_and_lists(c_rep, ap_col[off2], p_rep, off2, n_c, n_c)
    crow = np.repeat(np.arange(n_c, dtype=np.int32), np.diff(c_rowptr))
    c_diag = -np.ones(n_c, np.int32)
    dsel = np.flatnonzero(c_col == crow)
    c_diag[crow[dsel]] = dsel
    assert np.all(c_diag >= 0)
    # transpose map of P (restriction as a plain SpMV with R = P^T)
    order = np.lexsort((prow, p_col))
    pt_col = prow[order].astype(np.int32)
    pt_src = order.astype(np.int32)
    pt_rowptr = np.zeros(n_c + 1, np.int64)
    np.add.at(pt_rowptr, p_col + 1, 1)
    pt_rowptr = np.cumsum(pt_rowptr).astype(np.int32)
    # aggregate members (centroids)
    morder = np.argsort(agg, kind='stable').astype(np.int32)
    mem_ptr = np.zeros(n_c + 1, np.int64)
    np.add.at(mem_ptr, agg + 1, 1)
    mem_ptr = np.cumsum(mem_ptr).astype(np.int32)
    return dict(n_f=n_f, n_c=n_c, agg=agg.astype(np.int32), p_rowptr=p_rowptr, p_col=p_col, p_own=p_own,
                ps_ptr=ps_ptr, ps_a=ps_a, ps_j=ps_j, ap_rowptr=ap_rowptr, ap_col=ap_col, apl_ptr=apl_ptr,
                apl_a=apl_a, apl_p=apl_p, c_rowptr=c_rowptr, c_col=c_col, c_diag=c_diag, cl_ptr=cl_ptr,
                cl_p=cl_p, cl_ap=cl_ap, pt_rowptr=pt_rowptr, pt_col=pt_col, pt_src=pt_src,
                mem_ptr=mem_ptr, mem=morder, nnz_p=nnz_p, nnz_ap=nnz_ap, nnz_c=c_col.shape[0])


def build_hierarchy(rowptr, colidx, max_coarse_nodes=64, max_levels=12):
    """Coarsen until the coarsest level has <= max_coarse_nodes nodes (dense solve there)."""
    levels = []
    rp, ci = np.asarray(rowptr, np.int32), np.asarray(colidx, np.int32)
    while rp.shape[0] - 1 > max_coarse_nodes and len(levels) < max_levels:
        agg, n_c = aggregate(rp, ci)
        if n_c >= rp.shape[0] - 1:       # no coarsening possible (isolated nodes)
            break
        lv = build_level(rp, ci, agg, n_c)
        levels.append(lv)
        rp, ci = lv['c_rowptr'], lv['c_col']
    return levels
