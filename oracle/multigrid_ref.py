"""TEST INFRASTRUCTURE (not part of the product): SciPy / NumPy restatement of the smoothed-aggregation
multigrid preconditioner whose numeric part runs on the GPU (jaxsso_b200/csrc/jsso_multigrid.cuh), and a plain
Python restatement of the greedy aggregation (`jsso_mg_aggregate`).  The multigrid tier has no counterpart in
the reference (it solves with SuperLU, JaxSSO/solver.py:195-197); this file is what the tests compare the
product's symbolic gather lists, native aggregation and V-cycle quality against.  Only tests/ may import it."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def aggregate_py(rowptr, colidx):
    """Greedy aggregation (Vanek et al.): a node whose whole neighbourhood is free roots an
    aggregate of itself + neighbours; leftovers join the aggregate of their first aggregated
    neighbour (or become singletons).  Deterministic (ascending node order)."""
    n = rowptr.shape[0] - 1
    agg = -np.ones(n, np.int32)
    na = 0
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = colidx[rowptr[i]:rowptr[i + 1]]
        if np.all(agg[nb] < 0):
            agg[nb] = na
            agg[i] = na
            na += 1
    for i in range(n):
        if agg[i] < 0:
            nb = colidx[rowptr[i]:rowptr[i + 1]]
            got = agg[nb][agg[nb] >= 0]
            if got.size:
                agg[i] = got[0]
            else:
                agg[i] = na
                na += 1
    return agg, na


def rigid_blocks(X, cent, agg, mask_nodes=None):
    """T_i (n,6,6) about the aggregate centroids; rows of prescribed dofs zeroed."""
    r = X - cent[agg]
    T = np.zeros((X.shape[0], 6, 6))
    T[:, :3, :3] = np.eye(3)
    T[:, 3:, 3:] = np.eye(3)
    T[:, 0, 4] = r[:, 2]; T[:, 0, 5] = -r[:, 1]
    T[:, 1, 3] = -r[:, 2]; T[:, 1, 5] = r[:, 0]
    T[:, 2, 3] = r[:, 1]; T[:, 2, 4] = -r[:, 0]
    if mask_nodes is not None:
        T = T * (~mask_nodes)[:, :, None]
    return T


def centroids(X, lv):
    cnt = np.diff(lv['mem_ptr'])
    return np.stack([np.bincount(lv['agg'], weights=X[:, c], minlength=lv['n_c']) for c in range(3)], 1) / cnt[:, None]


class RefLevel:
    pass


def reference_setup(A0_bsr_blocks, rowptr, colidx, X0, mask_nodes, levels, Lt0=None, n_power=30):
    """SciPy reference of the numeric setup.  A0 blocks are (nnzb,6,6) [row, col]-oriented.
    ``Lt0``: optional (n,6,6) left factors applied to T at level 0 (L_i^T when the fine matrix is
    the block-Jacobi-scaled one).  Returns a list of RefLevel (A, Dinv, lam, P) + coarsest A."""
    out = []
    A = sp.bsr_matrix((A0_bsr_blocks, colidx, rowptr), shape=(6 * (rowptr.shape[0] - 1),) * 2).tocsr()
    X, mk, Lt = X0, mask_nodes, Lt0
    for lv in levels:
        n = lv['n_f']
        R = RefLevel()
        R.A = A
        Ab = A.tobsr((6, 6))
        D = np.zeros((n, 6, 6))
        for i in range(n):
            for k in range(Ab.indptr[i], Ab.indptr[i + 1]):
                if Ab.indices[k] == i:
                    D[i] = Ab.data[k]
        bad = np.abs(np.einsum('nii->ni', D)).min(1) < 1e-300
        D[bad] = np.eye(6)
        R.Dinv = np.linalg.inv(D)
        Dinv_sp = sp.bsr_matrix((R.Dinv, np.arange(n), np.arange(n + 1)), shape=A.shape)
        v = np.random.default_rng(0).uniform(-1, 1, 6 * n)
        v /= np.linalg.norm(v)
        lam = 1.0
        for _ in range(n_power):
            w = Dinv_sp @ (A @ v)
            lam = np.linalg.norm(w)
            v = w / lam
        R.lam = 1.15 * lam
        cent = centroids(X, lv)
        T = rigid_blocks(X, cent, lv['agg'], mk)
        if Lt is not None:
            T = np.einsum('nij,njk->nik', Lt, T)
        Tt = sp.bsr_matrix((T, lv['agg'], np.arange(n + 1)), shape=(6 * n, 6 * lv['n_c'])).tocsr()
        omega = 4.0 / (3.0 * R.lam)
        R.P = (Tt - omega * (Dinv_sp @ (A @ Tt))).tocsr()
        out.append(R)
        A = (R.P.T @ A @ R.P).tocsr()
        X, mk, Lt = cent, None, None
    return out, A


def reference_vcycle(ref_levels, A_coarse_dense_inv, b, deg=2, ratio=4.0):
    """V-cycle with Chebyshev(deg) pre/post smoothing on D^-1 A, eigenvalue interval [lam/ratio, lam]."""
    def cheb(R, rhs, x, zero_guess):
        lmax, lmin = R.lam, R.lam / ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        n = rhs.shape[0] // 6
        app = lambda v: np.einsum('nij,nj->ni', R.Dinv, v.reshape(n, 6)).ravel()
        r = app(rhs if zero_guess else rhs - R.A @ x)
        sigma = theta / delta
        rho = 1.0 / sigma
        d = r / theta
        for k in range(deg):
            x = x + d
            if k == deg - 1:
                break
            r = app(rhs - R.A @ x)
            rho_new = 1.0 / (2 * sigma - rho)
            d = rho_new * rho * d + (2 * rho_new / delta) * r
            rho = rho_new
        return x

    def rec(l, rhs):
        if l == len(ref_levels):
            return A_coarse_dense_inv @ rhs
        R = ref_levels[l]
        x = cheb(R, rhs, np.zeros_like(rhs), True)
        x = x + R.P @ rec(l + 1, R.P.T @ (rhs - R.A @ x))
        return cheb(R, rhs, x, False)

    return rec(0, b)
