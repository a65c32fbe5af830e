"""Mesh generators: the reference's own validation meshes and the synthetic
benchmark meshes of SURVEY.md section 8(d).  Pure NumPy, no device code.

Every generator returns a ``MeshData`` with exactly the frozen arrays that
``Model.model_ready`` produces in the reference (model.py:221-338):
``crds`` f64 (n_node,3), ``cnct_quads`` i32 (n_q,4), ``prop_quads`` f64 (n_q,5)
= t,E,nu,kx,ky, ``cnct_beams`` i32 (n_b,2), ``prop_beams`` f64 (n_b,6) =
E,G,Iy,Iz,J,A, ``known`` i32 (fixed dof ids, in add_support order), ``loads``
f64 (6 n_node,), plus ``design_nodes`` (the nodes whose z is a design variable
in the corresponding notebook).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class MeshData:
    crds: np.ndarray
    cnct_quads: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))
    prop_quads: np.ndarray = field(default_factory=lambda: np.zeros((0, 5)))
    cnct_beams: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    prop_beams: np.ndarray = field(default_factory=lambda: np.zeros((0, 6)))
    known: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    loads: np.ndarray = None
    design_nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))

    def __post_init__(self):
        self.crds = np.ascontiguousarray(self.crds, dtype=np.float64).reshape(-1, 3)
        self.cnct_quads = np.ascontiguousarray(self.cnct_quads, dtype=np.int32).reshape(-1, 4)
        self.prop_quads = np.ascontiguousarray(self.prop_quads, dtype=np.float64).reshape(-1, 5)
        self.cnct_beams = np.ascontiguousarray(self.cnct_beams, dtype=np.int32).reshape(-1, 2)
        self.prop_beams = np.ascontiguousarray(self.prop_beams, dtype=np.float64).reshape(-1, 6)
        self.known = np.ascontiguousarray(self.known, dtype=np.int32).ravel()
        if self.loads is None:
            self.loads = np.zeros(6 * self.crds.shape[0])
        self.loads = np.ascontiguousarray(self.loads, dtype=np.float64).ravel()
        self.design_nodes = np.asarray(self.design_nodes, dtype=np.int64).ravel()

    @property
    def n_node(self):
        return self.crds.shape[0]

    @property
    def ndof(self):
        return 6 * self.crds.shape[0]

    @property
    def n_quad(self):
        return self.cnct_quads.shape[0]

    @property
    def n_beam(self):
        return self.cnct_beams.shape[0]


def support_dofs(nodes, active):
    """Fixed dof ids for ``add_support(node, active)`` calls (model.py:183-200)."""
    act = np.where(np.asarray(active) == 1)[0]
    return (6 * np.asarray(nodes, np.int64)[:, None] + act[None, :]).ravel().astype(np.int32)


def grid_quads(n_col, n_row):
    """Connectivity of the reference's structured quad grids
    (Test/shells_fea_validation.ipynb cell 1): LL, LR, UR, UL per element."""
    e = np.arange(n_col * n_row, dtype=np.int64)
    ir = e // n_col
    n4 = e + ir + n_col + 1
    n3 = n4 + 1
    n2 = n3 - (n_col + 1)
    n1 = n2 - 1
    return np.stack([n1, n2, n3, n4], axis=1).astype(np.int32)


def barrel_arch():
    """19x19 MITC4 barrel arch of Test/shells_fea_validation.ipynb cells 1, 3
    (identical in Test/shells_ad_validation.ipynb)."""
    n_col = n_row = 19
    x_span = 19
    xs = np.tile(np.linspace(0, x_span, n_col + 1), n_row + 1)
    ys = np.tile(np.linspace(0, x_span, n_row + 1), (n_col + 1, 1)).T.reshape(-1)
    zs = 0.05 * (-(np.linspace(0, x_span, n_col + 1) - x_span / 2) ** 2 + (x_span / 2) ** 2)
    zs = np.tile(zs, n_row + 1)
    zs = np.where(xs == 0, 0, zs)
    zs = np.where(xs == x_span, 0, zs)
    n_node = (n_col + 1) * (n_row + 1)
    cnct = grid_quads(n_col, n_row)
    t, E, nu = 0.25, 24855578. * 1e-3, 0.2
    edge = (xs == 0) | (xs == x_span)
    design = np.where(~edge)[0]
    fixed = np.where(edge)[0]
    Q = 500 * 400 / n_node
    loads = np.zeros(6 * n_node)
    loads[6 * design + 2] = -Q
    prop = np.tile([t, E, nu, 1.0, 1.0], (cnct.shape[0], 1))
    return MeshData(crds=np.stack([xs, ys, zs], 1), cnct_quads=cnct, prop_quads=prop,
                    known=support_dofs(fixed, [1, 1, 1, 0, 0, 0]), loads=loads,
                    design_nodes=design)


_BEAM_SECTION = dict(E=1.999E+08, G=1.999E+08 / (2 * (1 + 0.3)), Iy=6.572e-05, Iz=3.301e-06,
                     J=6.572e-05 + 3.301e-06, A=4.265e-03)


def _beam_props(n, s=_BEAM_SECTION):
    return np.tile([s['E'], s['G'], s['Iy'], s['Iz'], s['J'], s['A']], (n, 1))


def beam_arch():
    """100-node parabolic beam-column arch of Test/beamcols_fea_validation.ipynb cell 1."""
    n_node, Q, rise, x_span = 100, 500, 5, 10
    x = np.linspace(0, x_span, n_node)
    z = -(rise / (x_span ** 2 / 4)) * ((x - x_span / 2) ** 2 - x_span ** 2 / 4)
    z[0] = 0
    z[-1] = 0
    design = np.arange(1, n_node - 1)
    cnct = np.stack([np.arange(n_node - 1), np.arange(1, n_node)], 1)
    loads = np.zeros(6 * n_node)
    loads[6 * design + 2] = -Q
    return MeshData(crds=np.stack([x, np.zeros(n_node), z], 1), cnct_beams=cnct,
                    prop_beams=_beam_props(n_node - 1),
                    known=support_dofs([0, n_node - 1], [1, 1, 1, 1, 0, 1]), loads=loads,
                    design_nodes=design)


def frames(n, m=100):
    """Multi-span arch ``f(n, m)`` of Test/Frames_speed.ipynb cell 1: m spans of
    n beam-columns each (n*m elements, n*m+1 nodes)."""
    Q, rise, x_span = 50000, 10, 30
    n_node = m * n + 1
    x = np.linspace(0, x_span * m, n_node)
    z = -(rise / (x_span ** 2 / 4)) * ((x % x_span - x_span / 2) ** 2 - x_span ** 2 / 4)
    idx = np.arange(n_node)
    design = idx[idx % n != 0]
    fixed = idx[idx % n == 0]
    cnct = np.stack([idx[:-1], idx[1:]], 1)
    loads = np.zeros(6 * n_node)
    loads[6 * design + 2] = -Q / (m * (n - 1))
    return MeshData(crds=np.stack([x, np.zeros(n_node), z], 1), cnct_beams=cnct,
                    prop_beams=_beam_props(n_node - 1),
                    known=support_dofs(fixed, [1, 1, 1, 1, 0, 1]), loads=loads,
                    design_nodes=design)


def plate(N, jitter=True, t=0.25, E=2.4855578e7, nu=0.2, M=None):
    """Synthetic N x M doubly-curved MITC4 cap of SURVEY.md 8(d) (M defaults to N).

    Nodes (i,j) at x=i, y=j (h=1), node id j*(N+1)+i, z = 0.1 N (1-xi^2)(1-eta^2);
    supports [1,1,1,0,0,0] on every boundary node, f_z = -1 on interior nodes.
    ``jitter``: interior nodes get x,y,z += U(-0.01,0.01) from
    default_rng(20240723), drawn in node order, x then y then z.
    """
    M = N if M is None else M
    i = np.tile(np.arange(N + 1), M + 1)
    j = np.repeat(np.arange(M + 1), N + 1)
    xi = 2.0 * i / N - 1.0
    eta = 2.0 * j / M - 1.0
    crds = np.stack([i.astype(float), j.astype(float),
                     0.1 * N * (1 - xi ** 2) * (1 - eta ** 2)], 1)
    boundary = (i == 0) | (i == N) | (j == 0) | (j == M)
    interior = np.where(~boundary)[0]
    if jitter:
        rng = np.random.default_rng(20240723)
        crds[interior] += rng.uniform(-0.01, 0.01, size=(interior.shape[0], 3))
    cnct = grid_quads(N, M)
    loads = np.zeros(6 * crds.shape[0])
    loads[6 * interior + 2] = -1.0
    prop = np.tile([t, E, nu, 1.0, 1.0], (cnct.shape[0], 1))
    return MeshData(crds=crds, cnct_quads=cnct, prop_quads=prop,
                    known=support_dofs(np.where(boundary)[0], [1, 1, 1, 0, 0, 0]),
                    loads=loads, design_nodes=interior)


def gridshell(n=224, k=0):
    """Synthetic beam-column gridshell design k of SURVEY.md 8(d): n x n nodes,
    beams along grid lines, section of Examples/Gridshell_Station_Shape.ipynb."""
    i = np.tile(np.arange(n), n)
    j = np.repeat(np.arange(n), n)
    xi = 2.0 * i / (n - 1) - 1.0
    eta = 2.0 * j / (n - 1) - 1.0
    rng = np.random.default_rng(1000 + k)
    a_k = 5 + 0.25 * k
    z = a_k * (1 - xi ** 2) * (1 - eta ** 2) + rng.uniform(-0.01, 0.01, size=n * n)
    boundary = (i == 0) | (i == n - 1) | (j == 0) | (j == n - 1)
    nid = np.arange(n * n).reshape(n, n)
    cnct = np.concatenate([np.stack([nid[:, :-1].ravel(), nid[:, 1:].ravel()], 1),
                           np.stack([nid[:-1, :].ravel(), nid[1:, :].ravel()], 1)])
    # the z perturbation (kept on the boundary too) guarantees that no member is exactly parallel
    # to global Y, where the reference's transformation degenerates (element.py:92-94)
    d = np.stack([i, j, z], 1)[cnct[:, 1]] - np.stack([i, j, z], 1)[cnct[:, 0]]
    assert (np.hypot(d[:, 0], d[:, 2]) / np.linalg.norm(d, axis=1)).min() > 1e-9
    b, h, E = 0.1, 0.2, 3.79e9
    Iy, Iz = b * h ** 3 / 12, h * b ** 3 / 12
    sec = dict(E=E, G=E / 2.6, Iy=Iy, Iz=Iz, J=Iy + Iz, A=b * h)
    interior = np.where(~boundary)[0]
    loads = np.zeros(6 * n * n)
    loads[6 * interior + 2] = 10.0
    return MeshData(crds=np.stack([i.astype(float), j.astype(float), z], 1), cnct_beams=cnct,
                    prop_beams=_beam_props(cnct.shape[0], sec),
                    known=support_dofs(np.where(boundary)[0], [1, 1, 1, 0, 0, 0]), loads=loads,
                    design_nodes=interior)


def mannheim_quad(data):
    """Mannheim Multihalle MITC4 mesh (Examples/Data/Mannheim_Quad) with the
    seeded initial shape of SURVEY.md 8(d) C2.  ``data`` is a dict with keys
    x, y, cnct (n_q,4), bc_nodes (as stored in tests/golden/mannheim_quad.npz)."""
    xs = np.asarray(data['x'], float) - np.min(data['x'])
    ys = np.asarray(data['y'], float) - np.min(data['y'])
    n_node = xs.shape[0]
    bc = np.asarray(data['bc_nodes'], np.int64)
    is_bc = np.zeros(n_node, bool)
    is_bc[bc] = True
    design = np.where(~is_bc)[0]
    zs = np.zeros(n_node)
    zs[design] = 0.5 + 0.01 * np.random.default_rng(2).uniform(0, 1, size=design.shape[0])
    D = np.hypot(xs[:, None] - xs[None, :], ys[:, None] - ys[None, :])
    R = 10.0
    B = np.where(D > R, 0, (1 / R) * (R - D))
    B = B / B.sum(1, keepdims=True)
    zf = zs.copy()
    zf[design] = (B @ zs)[design]
    cnct = np.asarray(data['cnct'], np.int32).reshape(-1, 4)
    loads = np.zeros(6 * n_node)
    loads[6 * design + 2] = -5000.0
    prop = np.tile([0.1, 1e10, 0.3, 1.0, 1.0], (cnct.shape[0], 1))
    return MeshData(crds=np.stack([xs, ys, zf], 1), cnct_quads=cnct, prop_quads=prop,
                    known=support_dofs(np.where(is_bc)[0], [1, 1, 1, 1, 1, 1]), loads=loads,
                    design_nodes=design)
