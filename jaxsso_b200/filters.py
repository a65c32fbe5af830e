"""Sparse radius ("hat") filter, the chain-rule operator every example applies either side of the
hot path.

The reference's notebooks build a DENSE n x n matrix
    B_ij = w_ij / sum_j w_ij,   w_ij = max(0, (R - d_ij) / R),   d_ij = XY-projected distance
(Examples/Shells_Mannheim_Multihalle_Shape.ipynb cells 10-13; Examples/shells_topo_shape.ipynb cells
8-12) and use ``B @ z`` on the way in and ``sens @ B`` on the way out.  At 1M nodes the dense matrix
is 8 TB; here B is built once as CSR with a k-d tree (it depends on the XY layout only, which shape
optimisation in z leaves unchanged) and applied on the GPU with a CSR SpMV (B and B^T).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree

from . import _native as nat


class HatFilter:
    def __init__(self, xy, R, device=0):
        xy = np.ascontiguousarray(np.asarray(xy, dtype=np.float64)[:, :2])
        n = xy.shape[0]
        tree = cKDTree(xy)
        D = tree.sparse_distance_matrix(tree, R, output_type='coo_matrix')   # all pairs with d <= R
        off = D.row != D.col
        w = np.where(D.data[off] > R, 0.0, (1.0 / R) * (R - D.data[off]))
        W = (sp.coo_matrix((w, (D.row[off], D.col[off])), shape=(n, n)).tocsr() +
             sp.identity(n, format='csr'))                                    # w_ii = (R - 0)/R = 1
        W.sum_duplicates()
        rs = np.asarray(W.sum(axis=1)).ravel()
        self.B = sp.diags(1.0 / rs) @ W
        self.B = self.B.tocsr()
        self.B.sort_indices()
        self.BT = self.B.T.tocsr()
        self.BT.sort_indices()
        self.n = n
        nat.lib().jsso_set_device(device)
        self._dev = {}
        for name, M in (('B', self.B), ('BT', self.BT)):
            self._dev[name] = (nat.DeviceArray.from_host(M.indptr.astype(np.int32)),
                               nat.DeviceArray.from_host(M.indices.astype(np.int32)),
                               nat.DeviceArray.from_host(M.data.astype(np.float64)))

    def _apply(self, name, x):
        rp, ci, v = self._dev[name]
        xd = x if isinstance(x, nat.DeviceArray) else nat.DeviceArray.from_host(np.asarray(x, np.float64))
        yd = nat.DeviceArray((self.n,))
        rc = nat.lib().jsso_csr_spmv(self.n, rp.ptr, ci.ptr, v.ptr, xd.ptr, yd.ptr, None)
        if rc:
            raise nat.JssoError(rc, nat.lib().jsso_last_error(None).decode())
        return yd if isinstance(x, nat.DeviceArray) else yd.download()

    def apply(self, x):
        """filtered = B @ x   (the way in: physical variables from design variables)"""
        return self._apply('B', x)

    def apply_T(self, g):
        """B^T @ g = g @ B   (the way out: chain rule on the sensitivities)"""
        return self._apply('BT', g)
