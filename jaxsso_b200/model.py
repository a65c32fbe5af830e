"""Host-side mirror of the reference's model builder (JaxSSO/model.py) on top of
the B200 hot path.

Same public names, argument meaning and freezing rules as the reference
(``add_node/add_beamcol/add_quad/add_support/add_nodal_load/model_ready/
select_solver/solve/strain_energy``, model.py:56-383), but ``model_ready`` also
runs the one-time symbolic pass (block-CSR pattern, contributor maps, boundary
mask) and ``solve`` runs fused Ke+assembly and the block-Jacobi PCG on the GPU
instead of BCOO assembly + SuperLU.  Arrays are NumPy, not jax.numpy: the
JAX-traceable surface lives in ``jaxsso_b200.jax_ffi`` (import-guarded).
"""
from __future__ import annotations

import numpy as np

from . import _native as nat


class Node:
    """A node (reference: model.py:15-32): tag and global coordinates."""

    def __init__(self, nodeTag, X, Y, Z):
        self.nodeTag = nodeTag
        self.X, self.Y, self.Z = X, Y, Z


class Model:
    """The FE model to be analysed (reference: model.py:34-383)."""

    def __init__(self, device=0):
        self.nodes = {}          # nodeTag -> [X, Y, Z]
        self.beamcols = {}       # eleTag -> (i, j, E, G, Iy, Iz, J, A)
        self.quads = {}          # eleTag -> (i, j, m, n, t, E, nu, kx_mod, ky_mod)
        self.known_indices = []  # prescribed dof ids, in add_support order
        self.f = {}              # nodeTag -> [fx, fy, fz, mx, my, mz]
        self.u = None
        self.device = device
        self._handle = None
        self._topology_key = None
        self.multigrid_min_nodes = 20000
        self.last_stats = None

    # ---- building (reference: model.py:56-214) ------------------------------------
    def add_node(self, nodeTag, X, Y, Z):
        self.nodes[nodeTag] = [X, Y, Z]

    def update_node(self, nodeTag, XYZ, value):
        try:
            self.nodes[nodeTag][XYZ] = value
        except KeyError:
            print("Node {} does not exist in the model".format(nodeTag))

    def add_beamcol(self, eleTag, i_nodeTag, j_nodeTag, E, G, Iy, Iz, J, A):
        self.beamcols[eleTag] = (i_nodeTag, j_nodeTag, E, G, Iy, Iz, J, A)

    def add_truss(self, eleTag, i_nodeTag, j_nodeTag, E, A):
        """Not implemented in the reference either (model.py:130-151)."""
        pass

    def add_quad(self, eleTag, i_nodeTag, j_nodeTag, m_nodeTag, n_nodeTag, t, E, nu, kx_mod=1.0, ky_mod=1.0):
        self.quads[eleTag] = (i_nodeTag, j_nodeTag, m_nodeTag, n_nodeTag, t, E, nu, kx_mod, ky_mod)

    def add_support(self, nodeTag, active_supports=(1, 1, 1, 1, 1, 1)):
        act = np.flatnonzero(np.asarray(active_supports, dtype=np.int32) == 1)
        self.known_indices.extend((6 * int(nodeTag) + act).tolist())

    def add_nodal_load(self, nodeTag, nodal_load=(0.0, 0.0, 0.0, 0.0, 0.0, 0.0)):
        self.f[nodeTag] = list(nodal_load)

    # ---- freezing (reference: model.py:221-338) ------------------------------------
    def model_ready(self):
        """Freeze dict-of-elements into arrays and (re)build the symbolic state when the
        connectivity or the supports changed.  Node tags are row indices (model.py:197,
        268), so they must be 0..n_node-1 in insertion order, as in the reference."""
        self.crds = self.get_node_crds()
        self.nodal_loads = self.get_loads()
        self.known_id, self.unknown_id = self.get_boundary_ids()
        self.ndof = self.get_dofs()
        self.n_beamcol = len(self.beamcols)
        self.cnct_beamcols = self.get_cnct_beamcols()
        self.prop_beamcols = self.get_beamcols_cross_prop()
        self.n_quad = len(self.quads)
        self.cnct_quads = self.get_cnct_quads()
        self.prop_quads = self.get_quads_cross_prop()
        key = (self.crds.shape[0], self.cnct_quads.tobytes(), self.cnct_beamcols.tobytes(),
               self.known_id.tobytes(), self.device)
        if key != self._topology_key:
            if self._handle is not None:
                self._handle.close()
            self._handle = nat.Handle(self.crds.shape[0], self.cnct_quads, self.cnct_beamcols,
                                      np.unique(self.known_id), device=self.device)
            # larger models get the smoothed-aggregation multigrid hierarchy (symbolic, once); the
            # solver picks it automatically from 20 000 nodes on, block-Jacobi CG below
            if self.crds.shape[0] >= self.multigrid_min_nodes:
                self._handle.mg_setup()
            self._topology_key = key

    @property
    def handle(self):
        if self._handle is None:
            self.model_ready()
        return self._handle

    def get_node_crds(self):
        return np.array(list(self.nodes.values()), dtype=np.float64).reshape(-1, 3)

    def get_loads(self):
        load = np.zeros(self.get_dofs())
        for node, val in self.f.items():
            load[node * 6:node * 6 + 6] = val
        return load

    def get_boundary_ids(self):
        known = np.array(self.known_indices, dtype=np.int32).ravel()
        mask = np.ones(6 * len(self.nodes), dtype=bool)
        mask[known] = False
        return known, np.flatnonzero(mask).astype(np.int32)

    def get_dofs(self):
        return 6 * len(self.nodes)

    def get_cnct_beamcols(self):
        return np.array([b[:2] for b in self.beamcols.values()], dtype=np.int32).reshape(-1, 2)

    def get_cnct_quads(self):
        return np.array([q[:4] for q in self.quads.values()], dtype=np.int32).reshape(-1, 4)

    def get_beamcols_cross_prop(self):
        return np.array([b[2:] for b in self.beamcols.values()], dtype=np.float64).reshape(-1, 6)

    def get_quads_cross_prop(self):
        return np.array([q[4:] for q in self.quads.values()], dtype=np.float64).reshape(-1, 5)

    # ---- solving (reference: model.py:340-383) --------------------------------------
    def select_solver(self, which_solver='b200', enforce_scipy_sparse=True):
        """The reference returns one of three (K_aug, f_aug) -> u_aug callables
        (model.py:340-356).  Here every valid choice is served by the B200 block-Jacobi
        PCG on the reduced SPD system; the callable takes (crds, prop_beamcols,
        prop_quads) -- the fused plug-in point of SSO_model.params_u (SSO_model.py:243-248)."""
        if which_solver not in ('dense', 'sparse', 'b200'):
            print("Please select the right solver: dense or sparse")
            return None

        def solver(crds, prop_beamcols, prop_quads, opts=None):
            return self._solve_arrays(crds, prop_beamcols, prop_quads, opts)
        return solver

    def _solve_arrays(self, crds, prop_beamcols, prop_quads, opts=None):
        h = self.handle
        # a solver that stops at its attainable accuracy still returns its best u (with a RuntimeWarning), as the
        # reference's direct solvers return whatever accuracy they reach
        val, u, _, _, _, fs, _ = h.value_and_grad_host(crds, prop_quads, prop_beamcols, self.nodal_loads,
                                                       want=(), opts=opts, allow_noconv=True)
        self.last_stats = fs.as_dict()
        return u

    def solve(self, which_solver='b200', enforce_scipy_sparse=True, rtol=1e-10):
        self.model_ready()
        solver = self.select_solver(which_solver, enforce_scipy_sparse)
        if solver is None:
            return
        self.u = solver(self.crds, self.prop_beamcols, self.prop_quads, nat.make_opts(rtol=rtol))

    def strain_energy(self):
        if self.u is not None:
            return 0.5 * self.nodal_loads @ self.u
        print("Model has not been analyzed yet.")
