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


# ----------------------------------------------------------------------------- distributed replay
def _dofs(ids):
    return (6 * np.asarray(ids, np.int64)[:, None] + np.arange(6)[None, :]).ravel()


class _EmulatedRanks:
    """N simulated ranks, each with its OWN full-length copy of every level vector, initialised to NaN.  A rank
    only ever writes its row range; everything else arrives through `exchange` / `allgather`, exactly the calls
    `mg_solve_dist` (csrc/jsso_api.cu) makes.  A product that reads an entry nobody delivered turns NaN, so a
    missing ghost in the plan (jaxsso_b200/dist_multigrid.py) fails the comparison with the global V-cycle."""

    def __init__(self, ref_levels, A_coarse_inv, plan, deg):
        self.ref, self.Ainv, self.plan, self.deg = ref_levels, A_coarse_inv, plan, deg
        self.n_rank, self.n_dist = plan['n_rank'], plan['n_dist']
        self.v = [dict() for _ in range(self.n_rank)]
        self.sizes = [R.A.shape[0] for R in ref_levels] + [A_coarse_inv.shape[0]]
        self.n_exchange = 0

    def vec(self, r, l, name):
        key = (l, name)
        if key not in self.v[r]:
            self.v[r][key] = np.full(self.sizes[l], np.nan)
        return self.v[r][key]

    def rng(self, l, r):
        b = self.plan['bounds'][l]
        return 6 * int(b[r]), 6 * int(b[r + 1])

    def exchange(self, l, name):
        need = self.plan['need'][l]
        for r in range(self.n_rank):
            for s in range(self.n_rank):
                if s != r and need[r][s].size:
                    d = _dofs(need[r][s])
                    self.vec(r, l, name)[d] = self.vec(s, l, name)[d]
        self.n_exchange += 1

    def allgather(self, l, name):
        for r in range(self.n_rank):
            for s in range(self.n_rank):
                a, b = self.rng(l, s)
                self.vec(r, l, name)[a:b] = self.vec(s, l, name)[a:b]

    def rows(self, M, l_rows, r, x):
        """rows [range of rank r at level l_rows] of M @ x"""
        a, b = self.rng(l_rows, r)
        return M[a:b] @ x

    def dot(self, l, n1, n2):
        tot = 0.0
        for r in range(self.n_rank):
            a, b = self.rng(l, r)
            tot += float(self.vec(r, l, n1)[a:b] @ self.vec(r, l, n2)[a:b])
        return tot

    # ---- V-cycle, mirroring mg_smooth_dist / mg_vcycle_dist
    def _dinv(self, l, r, v):
        a, b = self.rng(l, r)
        D = self.ref[l].Dinv[a // 6:b // 6]
        return np.einsum('nij,nj->ni', D, v.reshape(-1, 6)).ravel()

    def smooth(self, l, bname, xname, zero_guess):
        R = self.ref[l]
        lmax, lmin = R.lam, R.lam / 4.0
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        if not zero_guess:
            self.exchange(l, xname)
        for r in range(self.n_rank):
            a, b = self.rng(l, r)
            bv, x = self.vec(r, l, bname), self.vec(r, l, xname)
            res = bv[a:b] if zero_guess else bv[a:b] - self.rows(R.A, l, r, x)
            d = self._dinv(l, r, res) / theta
            self.vec(r, l, 'd')[a:b] = d
            x[a:b] = d if zero_guess else x[a:b] + d
        for _ in range(1, self.deg):
            self.exchange(l, xname)
            rho_new = 1.0 / (2.0 * sigma - rho)
            for r in range(self.n_rank):
                a, b = self.rng(l, r)
                bv, x, d = self.vec(r, l, bname), self.vec(r, l, xname), self.vec(r, l, 'd')
                res = self._dinv(l, r, bv[a:b] - self.rows(R.A, l, r, x))
                d[a:b] = rho_new * rho * d[a:b] + (2.0 * rho_new / delta) * res
                x[a:b] += d[a:b]
            rho = rho_new

    def vcycle(self, l, bname, xname):
        if l >= self.n_dist:       # replicated levels: every rank runs the plain V-cycle on its full copy
            for r in range(self.n_rank):
                b = self.vec(r, l, bname)
                self.vec(r, l, xname)[:] = reference_vcycle(self.ref[l:], self.Ainv, b, deg=self.deg)
            return
        R = self.ref[l]
        Pt = R.P.T.tocsr()
        self.smooth(l, bname, xname, True)
        self.exchange(l, xname)
        for r in range(self.n_rank):
            a, b = self.rng(l, r)
            self.vec(r, l, 'r')[a:b] = self.vec(r, l, bname)[a:b] - self.rows(R.A, l, r, self.vec(r, l, xname))
        self.exchange(l, 'r')
        for r in range(self.n_rank):
            a, b = self.rng(l + 1, r)
            self.vec(r, l + 1, 'b')[a:b] = self.rows(Pt, l + 1, r, self.vec(r, l, 'r'))
        if l + 1 == self.n_dist:
            self.allgather(l + 1, 'b')
        self.vcycle(l + 1, 'b', 'x')
        if l + 1 < self.n_dist:
            self.exchange(l + 1, 'x')
        for r in range(self.n_rank):
            a, b = self.rng(l, r)
            self.vec(r, l, xname)[a:b] += self.rows(R.P, l, r, self.vec(r, l + 1, 'x'))
        self.smooth(l, bname, xname, False)


def emulate_distributed_vcycle(ref_levels, A_coarse_inv, plan, b, deg=2):
    """z = V-cycle(b) computed by row ranges; returns (z assembled from the owners, number of exchanges)."""
    E = _EmulatedRanks(ref_levels, A_coarse_inv, plan, deg)
    for r in range(E.n_rank):
        a, e = E.rng(0, r)
        E.vec(r, 0, 'pr')[a:e] = b[a:e]
    E.vcycle(0, 'pr', 'pz')
    z = np.empty_like(b)
    for r in range(E.n_rank):
        a, e = E.rng(0, r)
        z[a:e] = E.vec(r, 0, 'pz')[a:e]
    return z, E.n_exchange


def emulate_distributed_pcg(ref_levels, A_coarse_inv, plan, b, deg=2, rtol=1e-8, maxiter=200):
    """The outer PCG of mg_solve_dist: range updates, one exchange of p per iteration, three summed dots."""
    E = _EmulatedRanks(ref_levels, A_coarse_inv, plan, deg)
    A = ref_levels[0].A
    for r in range(E.n_rank):
        a, e = E.rng(0, r)
        E.vec(r, 0, 'b')[a:e] = b[a:e]
        E.vec(r, 0, 'pr')[a:e] = b[a:e]
        E.vec(r, 0, 'px')[a:e] = 0.0
    bb = E.dot(0, 'b', 'b')
    rr, rz, it = E.dot(0, 'pr', 'pr'), 0.0, 0
    while np.sqrt(rr / bb) > rtol and it < maxiter:
        E.vcycle(0, 'pr', 'pz')
        rz_new = E.dot(0, 'pr', 'pz')
        for r in range(E.n_rank):
            a, e = E.rng(0, r)
            p = E.vec(r, 0, 'pp')
            p[a:e] = E.vec(r, 0, 'pz')[a:e] if it == 0 else E.vec(r, 0, 'pz')[a:e] + (rz_new / rz) * p[a:e]
        rz = rz_new
        E.exchange(0, 'pp')
        for r in range(E.n_rank):
            a, e = E.rng(0, r)
            E.vec(r, 0, 'pq')[a:e] = E.rows(A, 0, r, E.vec(r, 0, 'pp'))
        alpha = rz / E.dot(0, 'pp', 'pq')
        for r in range(E.n_rank):
            a, e = E.rng(0, r)
            E.vec(r, 0, 'px')[a:e] += alpha * E.vec(r, 0, 'pp')[a:e]
            E.vec(r, 0, 'pr')[a:e] -= alpha * E.vec(r, 0, 'pq')[a:e]
        rr = E.dot(0, 'pr', 'pr')
        it += 1
    E.allgather(0, 'px')
    xs = [E.vec(r, 0, 'px') for r in range(E.n_rank)]
    for x in xs[1:]:
        assert np.array_equal(x, xs[0])
    return xs[0].copy(), it
