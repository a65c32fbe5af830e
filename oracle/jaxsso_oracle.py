"""CPU oracle: a NumPy/SciPy FP64 restatement of JaxSSO's per-gradient hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product path (``jaxsso_b200``) never
does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference is pure Python on jax/jaxlib, which is
not installable in this image (no ``jax`` wheel, no network), so the oracle is
pinned against the outputs the reference itself stored in its ``Test/*.ipynb``
notebooks (``tests/golden/reference_golden.json``, extracted by
``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks every
one of them.

Every function cites the reference lines it restates (paths relative to the
reference checkout, ``JaxSSO/...``).  All element routines are batched over a
leading element axis and dtype-generic: run with complex inputs they give the
exact element derivative by complex-step differentiation (``abs`` is written as
``a*sign(Re a)``, norms as ``sqrt(sum v*v)``, ``min`` picks the first arg-min of
the real part -- the rule XLA's reverse mode applies when no two candidates are
bitwise equal).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

GP = 1.0 / 3.0 ** 0.5  # element.py:935, 1047


# ----------------------------------------------------------------------------
# dtype-generic helpers
# ----------------------------------------------------------------------------
def _norm(v):
    """Euclidean norm over the last axis, analytic for complex-step inputs."""
    return np.sqrt(np.sum(v * v, axis=-1))


def _abs(a):
    """|a| with the derivative sign(a) (complex-step safe)."""
    if np.iscomplexobj(a):
        return a * np.where(a.real < 0, -1.0, 1.0)
    return np.abs(a)


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


# ----------------------------------------------------------------------------
# Beam-column (element.py:66-149)
# ----------------------------------------------------------------------------
def beam_T(crds):
    """12x12 transformation of a beam-column; element.py:66-105.  crds (n,6)."""
    crds = np.asarray(crds)
    x1, y1, z1, x2, y2, z2 = (crds[:, k] for k in range(6))
    length = np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2 + (z1 - z2) ** 2)
    Cx = (x2 - x1) / length
    Cy = (y2 - y1) / length
    Cz = (z2 - z1) / length
    Cxz = np.sqrt(Cx * Cx + Cz * Cz)  # norm of (Cx, 0, Cz), element.py:83-84
    n = crds.shape[0]
    dc = np.zeros((n, 3, 3), dtype=crds.dtype)
    vert = (Cxz.real == 0)  # member parallel to global Y, element.py:92-94
    safe = np.where(vert, 1.0, Cxz)
    # general branch, element.py:95-97 with sin_alpha=0, cos_alpha=1
    dc[:, 0, 0], dc[:, 0, 1], dc[:, 0, 2] = Cx, Cy, Cz
    dc[:, 1, 0], dc[:, 1, 1], dc[:, 1, 2] = Cz / safe, 0.0, -Cx / safe
    dc[:, 2, 0], dc[:, 2, 1], dc[:, 2, 2] = -Cx * Cy / safe, Cxz, -Cy * Cz / safe
    if np.any(vert):
        v = np.where(vert)[0]
        dc[v] = 0.0
        dc[v, 0, 1] = Cy[v]
        dc[v, 1, 0] = -Cy[v]
        dc[v, 1, 2] = -1.0
        dc[v, 2, 0] = -Cy[v]
    T = np.zeros((n, 12, 12), dtype=crds.dtype)
    for b in range(4):
        T[:, 3 * b:3 * b + 3, 3 * b:3 * b + 3] = dc
    return T


def beam_K_local(crds, E, G, Iy, Iz, J, A):
    """12x12 local stiffness of a beam-column; element.py:107-128."""
    crds = np.asarray(crds)
    x1, y1, z1, x2, y2, z2 = (crds[:, k] for k in range(6))
    L = np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2 + (z1 - z2) ** 2)
    n = crds.shape[0]
    dt = np.result_type(crds.dtype, np.asarray(E).dtype, np.asarray(Iy).dtype,
                        np.asarray(Iz).dtype, np.asarray(A).dtype,
                        np.asarray(G).dtype, np.asarray(J).dtype)
    k = np.zeros((n, 12, 12), dtype=dt)
    ax = A * E / L
    tz = G * J / L
    bz12, bz6, bz4, bz2 = 12 * E * Iz / L ** 3, 6 * E * Iz / L ** 2, 4 * E * Iz / L, 2 * E * Iz / L
    by12, by6, by4, by2 = 12 * E * Iy / L ** 3, 6 * E * Iy / L ** 2, 4 * E * Iy / L, 2 * E * Iy / L
    # rows exactly as the literal matrix, element.py:115-126
    k[:, 0, 0], k[:, 0, 6] = ax, -ax
    k[:, 1, 1], k[:, 1, 5], k[:, 1, 7], k[:, 1, 11] = bz12, bz6, -bz12, bz6
    k[:, 2, 2], k[:, 2, 4], k[:, 2, 8], k[:, 2, 10] = by12, -by6, -by12, -by6
    k[:, 3, 3], k[:, 3, 9] = tz, -tz
    k[:, 4, 2], k[:, 4, 4], k[:, 4, 8], k[:, 4, 10] = -by6, by4, by6, by2
    k[:, 5, 1], k[:, 5, 5], k[:, 5, 7], k[:, 5, 11] = bz6, bz4, -bz6, bz2
    k[:, 6, 0], k[:, 6, 6] = -ax, ax
    k[:, 7, 1], k[:, 7, 5], k[:, 7, 7], k[:, 7, 11] = -bz12, -bz6, bz12, -bz6
    k[:, 8, 2], k[:, 8, 4], k[:, 8, 8], k[:, 8, 10] = -by12, by6, by12, by6
    k[:, 9, 3], k[:, 9, 9] = -tz, tz
    k[:, 10, 2], k[:, 10, 4], k[:, 10, 8], k[:, 10, 10] = -by6, by2, by6, by4
    k[:, 11, 1], k[:, 11, 5], k[:, 11, 7], k[:, 11, 11] = bz6, bz2, -bz6, bz4
    return k


def element_K_beamcol(crds, prop):
    """Global 12x12 K_e = solve(T, K_local) @ T; element.py:130-139.

    crds (n,6) = x1,y1,z1,x2,y2,z2; prop (n,6) = E,G,Iy,Iz,J,A (model.py:315-326).
    """
    prop = np.asarray(prop)
    kl = beam_K_local(crds, *(prop[:, k] for k in range(6)))
    T = beam_T(crds)
    dt = np.result_type(kl.dtype, T.dtype)
    return np.linalg.solve(T.astype(dt), kl.astype(dt)) @ T


def beam_indices(cnct):
    """COO (row, col) of every raw beam entry; element.py:141-149, 272-273."""
    cnct = np.asarray(cnct, dtype=np.int64)
    dof = (6 * cnct[:, :, None] + np.arange(6)[None, None, :]).reshape(-1, 12)
    rows = np.repeat(dof, 12, axis=1)
    cols = np.tile(dof, (1, 12))
    return rows.reshape(-1), cols.reshape(-1)


# ----------------------------------------------------------------------------
# MITC4 quad (element.py:488-1106)
# ----------------------------------------------------------------------------
def quad_area(crds):
    """Surface area of a quad as two triangles (1,2,4) and (3,4,2); element.py:471-487.  crds (n,12)."""
    P = np.asarray(crds).reshape(-1, 4, 3)
    a1 = _norm(_cross(P[:, 1] - P[:, 0], P[:, 3] - P[:, 0])) * 0.5
    a2 = _norm(_cross(P[:, 3] - P[:, 2], P[:, 1] - P[:, 2])) * 0.5
    return a1 + a2


def quad_frame(crds):
    """Local axes (x^, y^, z^) of a quad; element.py:502-521 and 649-671.

    crds (n,12) = X1,Y1,Z1,...,X4,Y4,Z4.  Returns dirCos (n,3,3), rows x^,y^,z^,
    and the vectors 3->1, 3->2, 3->4.
    """
    P = np.asarray(crds).reshape(-1, 4, 3)
    v31 = P[:, 0] - P[:, 2]
    v32 = P[:, 1] - P[:, 2]
    v34 = P[:, 3] - P[:, 2]
    v42 = P[:, 1] - P[:, 3]
    x_axis = v31
    z_axis = _cross(x_axis, v42)
    y_axis = _cross(z_axis, x_axis)
    x_axis = x_axis / _norm(x_axis)[:, None]
    y_axis = y_axis / _norm(y_axis)[:, None]
    z_axis = z_axis / _norm(z_axis)[:, None]
    return np.stack([x_axis, y_axis, z_axis], axis=1), v31, v32, v34


def quad_loc_crds(crds):
    """Projected local 2-D coordinates x1,y1,...,x4,y4; element.py:488-539."""
    dc, v31, v32, v34 = quad_frame(crds)
    xa, ya = dc[:, 0], dc[:, 1]
    zero = np.zeros(v31.shape[0], dtype=dc.dtype)
    return np.stack([np.sum(v31 * xa, -1), np.sum(v31 * ya, -1),
                     np.sum(v32 * xa, -1), np.sum(v32 * ya, -1),
                     zero, zero,
                     np.sum(v34 * xa, -1), np.sum(v34 * ya, -1)], axis=1)


def quad_T(crds):
    """24x24 block-diagonal transformation; element.py:643-696."""
    dc = quad_frame(crds)[0]
    T = np.zeros((dc.shape[0], 24, 24), dtype=dc.dtype)
    for b in range(8):
        T[:, 3 * b:3 * b + 3, 3 * b:3 * b + 3] = dc
    return T


def quad_J(xy, r, s):
    """2x2 Jacobian at (r,s); element.py:698-709.  xy (n,8)."""
    x1, y1, x2, y2, x3, y3, x4, y4 = (xy[:, k] for k in range(8))
    J = np.empty((xy.shape[0], 2, 2), dtype=xy.dtype)
    J[:, 0, 0] = x1 * (s + 1) - x2 * (s + 1) + x3 * (s - 1) - x4 * (s - 1)
    J[:, 0, 1] = y1 * (s + 1) - y2 * (s + 1) + y3 * (s - 1) - y4 * (s - 1)
    J[:, 1, 0] = x1 * (r + 1) - x2 * (r - 1) + x3 * (r - 1) - x4 * (r + 1)
    J[:, 1, 1] = y1 * (r + 1) - y2 * (r - 1) + y3 * (r - 1) - y4 * (r + 1)
    return 0.25 * J


def _dH(xy, r, s):
    """J^-1 dN, the (n,2,4) physical shape-function gradients; element.py:721-722, 811-812."""
    dN = 0.25 * np.array([[1 + s, -1 - s, -1 + s, 1 - s],
                          [1 + r, 1 - r, -1 + r, -1 - r]])
    J = quad_J(xy, r, s)
    return np.linalg.solve(J, np.broadcast_to(dN, (xy.shape[0], 2, 4)).astype(J.dtype))


def _det2(J):
    return J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]


def quad_B_kappa(xy, r, s):
    """Bending B (n,3,12); element.py:711-733."""
    dH = _dH(xy, r, s)
    B = np.zeros((xy.shape[0], 3, 12), dtype=dH.dtype)
    for k in range(4):
        B[:, 0, 3 * k + 2] = -dH[:, 0, k]
        B[:, 1, 3 * k + 1] = dH[:, 1, k]
        B[:, 2, 3 * k + 1] = dH[:, 0, k]
        B[:, 2, 3 * k + 2] = -dH[:, 1, k]
    return B


def quad_B_gamma_MITC4(xy, r, s):
    """MITC4 shear B (n,2,12); element.py:735-801."""
    x1, y1, x2, y2, x3, y3, x4, y4 = (xy[:, k] for k in range(8))
    Ax = x1 - x2 - x3 + x4
    Bx = x1 - x2 + x3 - x4
    Cx = x1 + x2 - x3 - x4
    Ay = y1 - y2 - y3 + y4
    By = y1 - y2 + y3 - y4
    Cy = y1 + y2 - y3 - y4
    rax = np.stack([(x1 + x4) / 2 - (x2 + x3) / 2, (y1 + y4) / 2 - (y2 + y3) / 2], -1)
    sax = np.stack([(x1 + x2) / 2 - (x3 + x4) / 2, (y1 + y2) / 2 - (y3 + y4) / 2], -1)
    rax = rax / _norm(rax)[:, None]
    sax = sax / _norm(sax)[:, None]
    det_J = _det2(quad_J(xy, r, s))
    gr = ((Cx + r * Bx) ** 2 + (Cy + r * By) ** 2) ** 0.5 / (8 * det_J)
    gs = ((Ax + s * Bx) ** 2 + (Ay + s * By) ** 2) ** 0.5 / (8 * det_J)
    one = np.ones_like(x1)
    grz = gr[:, None] * np.stack([
        (1 + s) / 2 * one, -(y1 - y2) * (1 + s) / 4, (x1 - x2) * (1 + s) / 4,
        -(1 + s) / 2 * one, -(y1 - y2) * (1 + s) / 4, (x1 - x2) * (1 + s) / 4,
        -(1 - s) / 2 * one, -(y4 - y3) * (1 - s) / 4, (x4 - x3) * (1 - s) / 4,
        (1 - s) / 2 * one, -(y4 - y3) * (1 - s) / 4, (x4 - x3) * (1 - s) / 4], axis=1)
    gsz = gs[:, None] * np.stack([
        (1 + r) / 2 * one, -(y1 - y4) * (1 + r) / 4, (x1 - x4) * (1 + r) / 4,
        (1 - r) / 2 * one, -(y2 - y3) * (1 - r) / 4, (x2 - x3) * (1 - r) / 4,
        -(1 - r) / 2 * one, -(y2 - y3) * (1 - r) / 4, (x2 - x3) * (1 - r) / 4,
        -(1 + r) / 2 * one, -(y1 - y4) * (1 + r) / 4, (x1 - x4) * (1 + r) / 4], axis=1)
    cos_alpha = rax[:, 0]
    cos_beta = sax[:, 0]
    # |cross(axis, e_x)| = |axis_y|; signs hard-wired as element.py:795-796
    sin_alpha = -_abs(rax[:, 1])
    sin_beta = _abs(sax[:, 1])
    return np.stack([grz * sin_beta[:, None] - gsz * sin_alpha[:, None],
                     -grz * cos_beta[:, None] + gsz * cos_alpha[:, None]], axis=1)


def quad_B_m(xy, r, s):
    """Membrane B (n,3,8); element.py:803-818."""
    dH = _dH(xy, r, s)
    B = np.zeros((xy.shape[0], 3, 8), dtype=dH.dtype)
    for k in range(4):
        B[:, 0, 2 * k] = dH[:, 0, k]
        B[:, 1, 2 * k + 1] = dH[:, 1, k]
        B[:, 2, 2 * k] = dH[:, 1, k]
        B[:, 2, 2 * k + 1] = dH[:, 0, k]
    return B


def quad_Cb(nu, E, h):
    """element.py:820-833."""
    z = np.zeros_like(nu * E * h)
    o = z + 1
    return (E * h ** 3 / (12 * (1 - nu ** 2)))[:, None, None] * np.stack(
        [np.stack([o, nu + z, z], -1), np.stack([nu + z, o, z], -1),
         np.stack([z, z, (1 - nu) / 2 + z], -1)], axis=1)


def quad_Cs(nu, E, h):
    """element.py:835-849."""
    c = E * h * (5 / 6) / (2 * (1 + nu))
    z = np.zeros_like(c)
    return np.stack([np.stack([c, z], -1), np.stack([z, c], -1)], axis=1)


def quad_Cm(nu, E, kx, ky):
    """element.py:851-875 (unsymmetric when kx != ky, as in the reference)."""
    Ex, Ey = E * kx, E * ky
    G = E / (2 * (1 + nu))
    z = np.zeros_like(Ex * Ey * nu)
    pre = 1 / (1 - nu * nu)
    return pre[:, None, None] * np.stack(
        [np.stack([Ex + z, nu * Ex + z, z], -1), np.stack([nu * Ey + z, Ey + z, z], -1),
         np.stack([z, z, (1 - nu * nu) * G + z], -1)], axis=1)


_GPS = ((GP, GP), (-GP, GP), (-GP, -GP), (GP, -GP))  # element.py:939-949
_KB_DOF = np.array([2, 3, 4, 8, 9, 10, 14, 15, 16, 20, 21, 22])  # element.py:895-919
_KM_DOF = np.array([0, 1, 6, 7, 12, 13, 18, 19])  # element.py:1014-1030
_KRZ_DIAG = np.array([1, 2, 4, 5, 7, 8, 10, 11])  # element.py:978


def quad_k_b(crds, t, E, nu, return_parts=False):
    """Expanded 24x24 bending+shear+drilling local stiffness; element.py:923-995."""
    xy = quad_loc_crds(crds)
    Cb, Cs = quad_Cb(nu, E, t), quad_Cs(nu, E, t)
    n = xy.shape[0]
    k1 = 0
    k2 = 0
    for (r, s) in _GPS:
        dJ = _det2(quad_J(xy, r, s))[:, None, None]
        Bk = quad_B_kappa(xy, r, s)
        Bg = quad_B_gamma_MITC4(xy, r, s)
        k1 = k1 + np.swapaxes(Bk, 1, 2) @ (Cb @ Bk) * dJ
        k2 = k2 + np.swapaxes(Bg, 1, 2) @ (Cs @ Bg) * dJ
    k = k1 + k2
    diag = _abs(k[:, _KRZ_DIAG, _KRZ_DIAG])
    imin = np.argmin(diag.real, axis=1)  # first arg-min (ties -> lowest index)
    k_rz = diag[np.arange(n), imin] / 1000
    kexp = np.zeros((n, 24, 24), dtype=k.dtype)
    kexp[:, _KB_DOF[:, None], _KB_DOF[None, :]] = k
    for d in (5, 11, 17, 23):
        kexp[:, d, d] = k_rz
    if return_parts:
        return kexp, k1, k2, k_rz, imin
    return kexp


def quad_k_m(crds, t, E, nu, kx, ky):
    """Expanded 24x24 membrane local stiffness; element.py:1035-1071."""
    xy = quad_loc_crds(crds)
    Cm = quad_Cm(nu, E, kx, ky)
    k = 0
    for (r, s) in _GPS:
        dJ = _det2(quad_J(xy, r, s))[:, None, None]
        B = quad_B_m(xy, r, s)
        k = k + np.swapaxes(B, 1, 2) @ (Cm @ B) * dJ
    k = t[:, None, None] * k
    kexp = np.zeros((xy.shape[0], 24, 24), dtype=k.dtype)
    kexp[:, _KM_DOF[:, None], _KM_DOF[None, :]] = k
    return kexp


def element_K_quad_local(crds, prop):
    """element.py:1086-1095.  prop (n,5) = t,E,nu,kx_mod,ky_mod (model.py:328-338)."""
    crds, prop = np.asarray(crds), np.asarray(prop)
    dt = np.result_type(crds.dtype, prop.dtype)
    crds, prop = crds.astype(dt), prop.astype(dt)
    t, E, nu, kx, ky = (prop[:, k] for k in range(5))
    return quad_k_m(crds, t, E, nu, kx, ky) + quad_k_b(crds, t, E, nu)  # k_b gets kx=ky=1


def element_K_quad(crds, prop):
    """Global 24x24 K_e = solve(T, K_local) @ T; element.py:1073-1084."""
    K = element_K_quad_local(crds, prop)
    T = quad_T(np.asarray(crds).astype(K.dtype))
    return np.linalg.solve(T, K) @ T


def quad_indices(cnct):
    """COO (row, col) of every raw quad entry; element.py:1097-1106, 1238-1239."""
    cnct = np.asarray(cnct, dtype=np.int64)
    dof = (6 * cnct[:, :, None] + np.arange(6)[None, None, :]).reshape(-1, 24)
    rows = np.repeat(dof, 24, axis=1)
    cols = np.tile(dof, (1, 24))
    return rows.reshape(-1), cols.reshape(-1)


# ----------------------------------------------------------------------------
# Global system (assemblemodel.py)
# ----------------------------------------------------------------------------
class Mesh:
    """Frozen model arrays, the output of Model.model_ready (model.py:221-246)."""

    def __init__(self, crds, cnct_quads=None, prop_quads=None, cnct_beams=None,
                 prop_beams=None, known=None, loads=None):
        self.crds = np.asarray(crds, dtype=float).reshape(-1, 3)
        self.n_node = self.crds.shape[0]
        self.ndof = 6 * self.n_node
        self.cnct_quads = (np.zeros((0, 4), np.int32) if cnct_quads is None
                           else np.asarray(cnct_quads, np.int32).reshape(-1, 4))
        self.prop_quads = (np.zeros((0, 5)) if prop_quads is None
                           else np.asarray(prop_quads, float).reshape(-1, 5))
        self.cnct_beams = (np.zeros((0, 2), np.int32) if cnct_beams is None
                           else np.asarray(cnct_beams, np.int32).reshape(-1, 2))
        self.prop_beams = (np.zeros((0, 6)) if prop_beams is None
                           else np.asarray(prop_beams, float).reshape(-1, 6))
        self.known = (np.zeros(0, np.int32) if known is None
                      else np.asarray(known, np.int32).ravel())
        self.loads = np.zeros(self.ndof) if loads is None else np.asarray(loads, float).ravel()
        self.n_quad = self.cnct_quads.shape[0]
        self.n_beam = self.cnct_beams.shape[0]

    def supports(self, nodes, active=(1, 1, 1, 1, 1, 1)):
        """model.py:183-200 (add_support), in call order."""
        act = np.where(np.asarray(active) == 1)[0]
        new = (6 * np.asarray(nodes, np.int64)[:, None] + act[None, :]).ravel()
        self.known = np.concatenate([self.known, new.astype(np.int32)])
        return self


def raw_coo(mesh, crds=None, prop_quads=None, prop_beams=None):
    """Un-reduced COO of K in the reference's order: [(0,0):0] ++ beams ++ quads.

    assemblemodel.py:196-213 (K_func); BCOO ``+`` concatenates its operands.
    """
    crds = mesh.crds if crds is None else crds
    pq = mesh.prop_quads if prop_quads is None else prop_quads
    pb = mesh.prop_beams if prop_beams is None else prop_beams
    rows, cols, data = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [np.zeros(1)]
    if mesh.n_beam > 0:
        e = np.asarray(crds)[mesh.cnct_beams].reshape(-1, 6)
        r, c = beam_indices(mesh.cnct_beams)
        rows.append(r), cols.append(c), data.append(element_K_beamcol(e, pb).reshape(-1))
    if mesh.n_quad > 0:
        e = np.asarray(crds)[mesh.cnct_quads].reshape(-1, 12)
        r, c = quad_indices(mesh.cnct_quads)
        rows.append(r), cols.append(c), data.append(element_K_quad(e, pq).reshape(-1))
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(data)


def K_global(mesh, **kw):
    """Sorted, duplicate-summed K as CSR (what sort_indices+sum_duplicates leave)."""
    r, c, d = raw_coo(mesh, **kw)
    return sp.coo_matrix((d, (r, c)), shape=(mesh.ndof, mesh.ndof)).tocsr()


def sorted_unique_pattern(mesh):
    """The (row, col) set of K after sort_indices + sum_duplicates
    (assemblemodel.py:160-162): lexicographically sorted unique pairs."""
    r, c, _ = raw_coo_indices_only(mesh)
    key = np.unique(r * np.int64(mesh.ndof) + c)
    return key // mesh.ndof, key % mesh.ndof


def raw_coo_indices_only(mesh):
    rows, cols = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    if mesh.n_beam > 0:
        r, c = beam_indices(mesh.cnct_beams)
        rows.append(r), cols.append(c)
    if mesh.n_quad > 0:
        r, c = quad_indices(mesh.cnct_quads)
        rows.append(r), cols.append(c)
    return np.concatenate(rows), np.concatenate(cols), None


def K_aug(mesh, K=None, **kw):
    """[[K, V^T],[V, 0]] with V[i, known[i]] = 1; assemblemodel.py:111-163."""
    K = K_global(mesh, **kw) if K is None else K
    nc = mesh.known.shape[0]
    V = sp.coo_matrix((np.ones(nc), (np.arange(nc), mesh.known)), shape=(nc, mesh.ndof))
    return sp.bmat([[K, V.T], [V, None]], format='csr')


def f_aug(mesh):
    """assemblemodel.py:166-194."""
    return np.concatenate([mesh.loads, np.zeros(mesh.known.shape[0])])


# ----------------------------------------------------------------------------
# Solvers (solver.py)
# ----------------------------------------------------------------------------
def solve_literal(mesh, K=None, rhs=None, transpose=False, **kw):
    """u_aug = spsolve(K_aug, f_aug): the reference's default path, solver.py:176-210."""
    A = K_aug(mesh, K=K, **kw)
    if transpose:
        A = A.T.tocsr()
    b = f_aug(mesh) if rhs is None else np.concatenate([rhs, np.zeros(mesh.known.shape[0])])
    return spla.spsolve(A.tocsc(), b)


def free_dofs(mesh):
    mask = np.ones(mesh.ndof, bool)
    mask[mesh.known] = False
    return np.where(mask)[0]


def solve_refined(mesh, K=None, rhs=None, transpose=False, steps=3, **kw):
    """Reduced SPD solve K_ff u_f = f_f (u_known = 0) + iterative refinement.

    Mathematically the same u as the Lagrange system of assemblemodel.py:111-163
    with zero prescribed displacements (assemblemodel.py:192); this is the
    accuracy reference for u / compliance / gradients (the literal augmented
    SuperLU solve is only good to ~1e-8)."""
    K = K_global(mesh, **kw) if K is None else K
    if transpose:
        K = K.T.tocsr()
    free = free_dofs(mesh)
    Kff = K[free][:, free].tocsc()
    b = (mesh.loads if rhs is None else rhs)[free]
    lu = spla.splu(Kff)
    x = lu.solve(b)
    for _ in range(steps):
        r = b - Kff @ x
        x = x + lu.solve(np.asarray(r, dtype=np.float64))
    u = np.zeros(mesh.ndof)
    u[free] = x
    return u


# ----------------------------------------------------------------------------
# Adjoint sensitivity (solver.py:138-166, 221-248 + XLA reverse mode of the rest)
# ----------------------------------------------------------------------------
def element_sensitivity(mesh, u, lam, crds=None, prop_quads=None, prop_beams=None,
                        h=1e-30):
    """sum_e sum_ab W_e[a,b] dK_e[a,b]/dx with W_e = -lam_e u_e^T.

    This is the VJP of -(K u - f) w.r.t. the COO data at cotangent lam
    (solver.py:157-166, 239-248) pulled back through vmap(element_K_*)
    (element.py:270, 1236) and the coordinate gathers (element.py:266-268,
    1229-1234).  dK_e/dx is exact (complex step, h = 1e-30).
    Returns d_crds (n_node,3), d_prop_quads (n_q,5), d_prop_beams (n_b,6).
    """
    crds = mesh.crds if crds is None else np.asarray(crds, float)
    pq = mesh.prop_quads if prop_quads is None else np.asarray(prop_quads, float)
    pb = mesh.prop_beams if prop_beams is None else np.asarray(prop_beams, float)
    d_crds = np.zeros((mesh.n_node, 3))
    d_pq = np.zeros((mesh.n_quad, 5))
    d_pb = np.zeros((mesh.n_beam, 6))

    def run(cnct, nn, prop, fn, d_prop):
        ne = cnct.shape[0]
        dof = (6 * cnct.astype(np.int64)[:, :, None] + np.arange(6)[None, None, :]).reshape(ne, -1)
        W = -lam[dof][:, :, None] * u[dof][:, None, :]
        e = crds[cnct].reshape(ne, 3 * nn)
        for k in range(3 * nn):
            ec = e.astype(complex)
            ec[:, k] += 1j * h
            dK = fn(ec, prop).imag / h
            np.add.at(d_crds, (cnct[:, k // 3], k % 3), np.sum(W * dK, axis=(1, 2)))
        for k in range(prop.shape[1]):
            pc = prop.astype(complex)
            pc[:, k] += 1j * h
            dK = fn(e, pc).imag / h
            d_prop[:, k] = np.sum(W * dK, axis=(1, 2))

    if mesh.n_beam > 0:
        run(mesh.cnct_beams, 2, pb, element_K_beamcol, d_pb)
    if mesh.n_quad > 0:
        run(mesh.cnct_quads, 4, pq, element_K_quad, d_pq)
    return d_crds, d_pq, d_pb


def value_and_grad(mesh, g_fn=None, literal=False, **kw):
    """One objective + gradient evaluation, SSO_model.py:313-339.

    Objective defaults to the strain energy 0.5 f.u (SSO_model.py:297-301), for
    which g = dL/du = f/2.  ``literal=True`` follows the reference to the letter
    (augmented SuperLU solves, solver.py:195-197 and :236); otherwise the refined
    reduced solves are used.  Returns (value, u, lam, d_crds, d_prop_q, d_prop_b).
    """
    K = K_global(mesh, **kw)
    solve = solve_literal if literal else solve_refined
    u = solve(mesh, K=K)[:mesh.ndof]
    if g_fn is None:
        value, g = 0.5 * mesh.loads @ u, 0.5 * mesh.loads
    else:
        value, g = g_fn(u)
    lam = solve(mesh, K=K, rhs=g, transpose=True)[:mesh.ndof]
    d = element_sensitivity(mesh, u, lam, **kw)
    return (value, u, lam) + d
