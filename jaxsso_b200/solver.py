"""Solver plug-in with the reference's callable signature ``(K_aug, f_aug) -> u_aug``
(JaxSSO/solver.py:75, 103, 177; selected by Model.select_solver, model.py:340-356), served by the
B200 CG / multigrid solvers.

The reference hands its solvers the Lagrange-augmented matrix [[K, V^T], [V, 0]] with
V[i, known[i]] = 1 (assemblemodel.py:111-163), which is indefinite.  This plug-in recovers
``known`` from the V block, solves the equivalent reduced SPD system K_ff u_f = f_f (u_known = 0,
assemblemodel.py:192) on the GPU and rebuilds the multipliers mu = (f - K u)[known], so that
unmodified reference-side code (and user objectives that read u_aug) gets the same vector.

NumPy/SciPy in, NumPy out; with JAX it is wrapped in ``jax.pure_callback`` exactly like the
reference's ``sci_sparse_solve`` (see jaxsso_b200.jax_ffi).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _native as nat

_cache = {}


def split_augmented(K_aug):
    """K_aug (any SciPy sparse / (data, indices, indptr, shape) CSR) -> (K csr, known dof ids)."""
    if isinstance(K_aug, tuple):
        data, indices, indptr, shape = K_aug
        K_aug = sp.csr_matrix((np.asarray(data), np.asarray(indices), np.asarray(indptr)), shape=tuple(shape))
    A = sp.csr_matrix(K_aug)
    A.sum_duplicates()
    n_aug = A.shape[0]
    # constraint rows: the trailing rows whose only non-zero is a single 1 (explicit zeros, e.g. the
    # reference's zero_BCOO entry, are ignored)
    n_cons = 0
    for r in range(n_aug - 1, -1, -1):
        v = A.data[A.indptr[r]:A.indptr[r + 1]]
        nzv = v[v != 0.0]
        if nzv.shape[0] == 1 and nzv[0] == 1.0:
            n_cons += 1
        else:
            break
    ndof = n_aug - n_cons
    if ndof % 6:
        raise ValueError(f'cannot split K_aug: {ndof} structural dofs is not a multiple of 6')
    V = A[ndof:, :ndof].tocsr()
    V.eliminate_zeros()
    known = V.indices[np.argsort(np.repeat(np.arange(n_cons), np.diff(V.indptr)), kind='stable')].astype(np.int32)
    return A[:ndof, :ndof].tocsr(), known


def b200_solve(K_aug, f_aug, rtol=1e-10, device=0):
    """Drop-in for ``solver.sci_sparse_solve`` / ``jax_sparse_solve`` / ``jax_dense_solve``."""
    f_aug = np.asarray(f_aug, dtype=np.float64)
    K, known = split_augmented(K_aug)
    ndof = K.shape[0]
    Kb = K.tobsr((6, 6))
    Kb.sort_indices()
    # the reference's pattern always holds the diagonal blocks of nodes that belong to an element;
    # add missing ones (isolated, fully fixed nodes) so that the identity rows have a home
    n = ndof // 6
    Kb = (Kb + sp.bsr_matrix((np.zeros((n, 6, 6)), np.arange(n), np.arange(n + 1)), shape=Kb.shape)).tobsr((6, 6))
    Kb.sort_indices()
    key = (n, Kb.indptr.tobytes(), Kb.indices.tobytes(), np.unique(known).tobytes(), device)
    h = _cache.get(key)
    if h is None:
        _cache.clear()      # one cached pattern (optimiser loops reuse it; a new model replaces it)
        h = nat.Handle.from_bsr(Kb.indptr, Kb.indices, np.unique(known), device=device)
        _cache[key] = h
    h.set_values(Kb.data, apply_bc=True)
    b = nat.DeviceArray.from_host(f_aug[:ndof])
    x = nat.DeviceArray((ndof,))
    h.pcg(b, x, opts=nat.make_opts(rtol=rtol))
    u = x.download()
    mu = (f_aug[:ndof] - K @ u)[known]
    return np.concatenate([u, mu])
