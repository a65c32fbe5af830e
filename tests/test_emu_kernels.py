"""Kernel LOGIC on the CPU: the product's CUDA kernel headers compiled by g++ against a SIMT emulator
(tests/emu/cuda_emu.h: one std::thread per CUDA thread, barriers for __syncthreads / warp shuffles) and compared
with SciPy.  This is test infrastructure -- the package never loads the emulated library, and it proves nothing
about speed; it exists so that indexing / layout / reduction-order bugs show up here, where there is no GPU.
Covered: the block-CSR SpMV family (FP64, FP32 and binary16 block storage, long rows, short-row kernel, row
RANGES with offset pointers as the distributed multigrid solve uses them), the halo pack / unpack pair, the
deterministic dot product and the Chebyshev smoother step."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from jaxsso_b200 import meshes, multigrid as mg
from oracle import multigrid_ref as mgref
from tests.conftest import ROOT
from tests.test_multigrid import scaled_system

sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))


@pytest.fixture(scope='module')
def emu():
    import build_emu
    L = C.CDLL(build_emu.build())
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.emu_spmv_axpby.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, i32]
    L.emu_spmv_short.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp]
    L.emu_spmv_main.argtypes = [i32, vp, vp, vp, vp, vp, i32]
    L.emu_to_float.argtypes = [i64, vp, vp]
    L.emu_to_half.argtypes = [i64, vp, vp]
    L.emu_halo_pack.argtypes = [i32, vp, vp, vp]
    L.emu_halo_unpack.argtypes = [i32, vp, vp, vp]
    L.emu_dot.argtypes = [i64, vp, vp, i32]
    L.emu_dot.restype = dbl
    L.emu_cheb.argtypes = [i32, i32, vp, vp, vp, vp, dbl, dbl, i32]
    return L


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hierarchy(n=10):
    md = meshes.plate(n)
    Ah, L, mask = scaled_system(md)
    levels = mg.build_hierarchy(Ah.indptr.astype(np.int32), Ah.indices.astype(np.int32), max_coarse_nodes=4)
    ref, Ac = mgref.reference_setup(Ah.data, Ah.indptr, Ah.indices, md.crds, mask, levels, Lt0=L.transpose(0, 2, 1))
    return Ah, levels, ref


def bsr_arrays(M):
    """scipy matrix -> (rowptr, colidx, blocks[row, col]) with sorted columns"""
    B = M.tobsr((6, 6))
    B.sort_indices()
    return B.indptr.astype(np.int32), B.indices.astype(np.int32), np.ascontiguousarray(B.data)


def colmajor(blocks):
    return np.ascontiguousarray(blocks.transpose(0, 2, 1)).ravel()    # vals[36 s + 6 j + i] = A_s[i, j]


@pytest.mark.parametrize('vt', [0, 1, 2])
@pytest.mark.parametrize('mode', [0, 2, 3])
def test_level_spmv_all_storage_types(emu, vt, mode):
    Ah, levels, ref = hierarchy(10)
    rng = np.random.default_rng(vt * 10 + mode)
    for lvl, M in enumerate([ref[0].A, (ref[0].A @ ref[0].A).tocsr()]):       # A^2: rows with up to 25 blocks
        rp, ci, blk = bsr_arrays(M)
        if vt == 2 and lvl == 1:
            blk = blk / np.abs(blk).max()                  # binary16 is only used where |a| <= 1
            M = sp.bsr_matrix((blk, ci, rp), shape=M.shape)
        n = rp.shape[0] - 1
        assert lvl == 0 or np.diff(rp).max() > 10
        v64 = colmajor(blk)
        if vt == 0:
            vals, tol = v64, 1e-14
        elif vt == 1:
            vals = np.empty(v64.shape[0], np.float32)
            emu.emu_to_float(v64.shape[0], P(v64), P(vals))
            tol = 3e-7
        else:
            vals = np.empty(v64.shape[0], np.float16)
            emu.emu_to_half(v64.shape[0], P(v64), P(vals))
            tol = 2e-3
        x, b, y0 = rng.standard_normal(6 * n), rng.standard_normal(6 * n), rng.standard_normal(6 * n)
        y = y0.copy()
        assert emu.emu_spmv_axpby(mode, vt, n, P(rp), P(ci), P(vals), P(x), P(y), P(b), 2) == 0
        Ax = M @ x
        want = {0: Ax, 2: b - Ax, 3: y0 + Ax}[mode]
        scale = np.abs(M).dot(np.abs(x)).max()
        assert np.abs(y - want).max() <= tol * scale
        if vt == 2:       # the rounded matrix itself is reproduced to FP64 accuracy
            q = blk.astype(np.float16).astype(np.float64)
            Aq = sp.bsr_matrix((q, ci, rp), shape=M.shape) @ x
            assert np.abs(y - {0: Aq, 2: b - Aq, 3: y0 + Aq}[mode]).max() <= 1e-13 * scale


@pytest.mark.parametrize('vt', [0, 1, 2])
def test_row_range_with_offset_pointers(emu, vt):
    """What mg_solve_dist does: rows [s, e) of A x with `rowptr + s`, `y + 6 s`, `b + 6 s`, full-length x."""
    Ah, levels, ref = hierarchy(10)
    rp, ci, blk = bsr_arrays(ref[0].A)
    n = rp.shape[0] - 1
    v64 = colmajor(blk)
    vals = v64
    if vt == 1:
        vals = np.empty(v64.shape[0], np.float32); emu.emu_to_float(v64.shape[0], P(v64), P(vals))
    if vt == 2:
        vals = np.empty(v64.shape[0], np.float16); emu.emu_to_half(v64.shape[0], P(v64), P(vals))
    rng = np.random.default_rng(7)
    x, b = rng.standard_normal(6 * n), rng.standard_normal(6 * n)
    s, e = 37, 95
    y = np.full(6 * n, np.nan)
    rp_off = rp[s:]                     # a view: the kernel sees rowptr + s, absolute slots
    assert rp_off.ctypes.data == rp.ctypes.data + 4 * s
    ys, bs = y[6 * s:], b[6 * s:]
    assert emu.emu_spmv_axpby(2, vt, e - s, P(rp_off), P(ci), P(vals), P(x), P(ys), P(bs), 1) == 0
    want = b - ref[0].A @ x
    tol = {0: 1e-13, 1: 3e-6, 2: 5e-3}[vt]
    assert np.abs(y[6 * s:6 * e] - want[6 * s:6 * e]).max() <= tol * np.abs(want).max()
    assert np.isnan(y[:6 * s]).all() and np.isnan(y[6 * e:]).all()      # nothing outside the range is written


@pytest.mark.parametrize('mode', [0, 3])
def test_short_row_kernel_on_prolongator_and_restrictor(emu, mode):
    Ah, levels, ref = hierarchy(10)
    rng = np.random.default_rng(3)
    Pm = ref[0].P
    for M in (Pm, Pm.T.tocsr()):
        B = M.tobsr((6, 6)); B.sort_indices()
        rp, ci, blk = B.indptr.astype(np.int32), B.indices.astype(np.int32), np.ascontiguousarray(B.data)
        v64 = colmajor(blk)
        v32 = np.empty(v64.shape[0], np.float32)
        emu.emu_to_float(v64.shape[0], P(v64), P(v32))
        nr, nc = M.shape[0] // 6, M.shape[1] // 6
        x, y0 = rng.standard_normal(6 * nc), rng.standard_normal(6 * nr)
        y = y0.copy()
        assert emu.emu_spmv_short(mode, nr, P(rp), P(ci), P(v32), P(x), P(y), None) == 0
        want = (M @ x) if mode == 0 else y0 + M @ x
        assert np.abs(y - want).max() <= 3e-7 * np.abs(M).dot(np.abs(x)).max() + 1e-12
        # the long-row FP32 kernel gives the same on rectangular matrices
        y2 = y0.copy()
        assert emu.emu_spmv_axpby(mode, 1, nr, P(rp), P(ci), P(v32), P(x), P(y2), None, 1) == 0
        assert np.abs(y2 - y).max() <= 1e-12 * np.abs(want).max()


def test_cg_spmv_kernel(emu):
    Ah, levels, ref = hierarchy(8)
    rp, ci, blk = bsr_arrays(ref[0].A)
    n = rp.shape[0] - 1
    v64 = colmajor(blk)
    x = np.random.default_rng(1).standard_normal(6 * n)
    y = np.zeros(6 * n)
    emu.emu_spmv_main(n, P(rp), P(ci), P(v64), P(x), P(y), 3)
    assert np.abs(y - ref[0].A @ x).max() <= 1e-13 * np.abs(y).max()


def test_halo_pack_unpack_roundtrip(emu):
    rng = np.random.default_rng(2)
    n = 300
    v = rng.standard_normal(6 * n)
    idx = rng.permutation(n)[:77].astype(np.int32)
    buf = np.zeros(6 * 77)
    emu.emu_halo_pack(77, P(idx), P(v), P(buf))
    assert np.array_equal(buf.reshape(-1, 6), v.reshape(-1, 6)[idx])
    w = np.full(6 * n, np.nan)
    emu.emu_halo_unpack(77, P(idx), P(buf), P(w))
    assert np.array_equal(w.reshape(-1, 6)[idx], v.reshape(-1, 6)[idx])
    rest = np.setdiff1d(np.arange(n), idx)
    assert np.isnan(w.reshape(-1, 6)[rest]).all()


def test_dot_and_chebyshev_step(emu):
    rng = np.random.default_rng(4)
    n = 1000
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    for grid in (1, 3):
        assert abs(emu.emu_dot(n, P(a), P(b), grid) - a @ b) <= 1e-12 * np.abs(a * b).sum()
    assert emu.emu_dot(0, P(a), P(b), 1) == 0.0            # an empty row range contributes exactly zero
    nn = 50
    Dinv = rng.standard_normal((nn, 6, 6))
    r, d, x = rng.standard_normal(6 * nn), rng.standard_normal(6 * nn), rng.standard_normal(6 * nn)
    t = np.einsum('nij,nj->ni', Dinv, r.reshape(nn, 6)).ravel()
    d1, x1 = d.copy(), x.copy()
    emu.emu_cheb(1, nn, P(Dinv), P(r), P(d1), P(x1), 0.0, 0.4, 1)
    assert np.allclose(d1, 0.4 * t) and np.allclose(x1, 0.4 * t)
    d2, x2 = d.copy(), x.copy()
    emu.emu_cheb(0, nn, P(Dinv), P(r), P(d2), P(x2), 0.3, 0.7, 0)
    assert np.allclose(d2, 0.3 * d + 0.7 * t) and np.allclose(x2, x + 0.3 * d + 0.7 * t)
    d3, x3 = d.copy(), x.copy()     # no Dinv: the fine level (unit diagonal blocks)
    emu.emu_cheb(1, nn, None, P(r), P(d3), P(x3), 0.0, 0.5, 0)
    assert np.allclose(d3, 0.5 * r) and np.allclose(x3, x + 0.5 * r)


# ------------------------------------------------------------------------------- Ke + assembly, adjoint
from oracle import jaxsso_oracle as orc          # noqa: E402
from tests.conftest import to_oracle_mesh       # noqa: E402


@pytest.fixture(scope='module')
def emu2(emu):
    vp, i32 = C.c_void_p, C.c_int32
    emu.emu_assemble.argtypes = [i32, i32, i32, vp, i32, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32]
    emu.emu_adjoint.argtypes = [i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]
    emu.emu_pattern.argtypes = [i32, i32, vp, i32, vp, vp, vp, vp]
    return emu


def emu_K(emu, md, path, apply_bc=False, task_ctas=2):
    nnzb = C.c_int64()
    assert emu.emu_pattern(md.n_node, md.n_quad, P(md.cnct_quads), md.n_beam, P(md.cnct_beams), C.byref(nnzb), None, None) == 0
    rp, ci = np.empty(md.n_node + 1, np.int32), np.empty(nnzb.value, np.int32)
    emu.emu_pattern(md.n_node, md.n_quad, P(md.cnct_quads), md.n_beam, P(md.cnct_beams), C.byref(nnzb), P(rp), P(ci))
    vals = np.full(36 * nnzb.value, np.nan)
    flags = C.c_int32()
    rc = emu.emu_assemble(path, md.n_node, md.n_quad, P(md.cnct_quads), md.n_beam, P(md.cnct_beams), md.known.shape[0],
                          P(md.known), P(md.crds), P(md.prop_quads), P(md.prop_beams), int(apply_bc), P(vals),
                          C.byref(flags), task_ctas)
    assert rc == 0, rc
    blocks = vals.reshape(-1, 6, 6).transpose(0, 2, 1)          # column-major storage -> [row, col]
    return sp.bsr_matrix((np.ascontiguousarray(blocks), ci, rp), shape=(md.ndof, md.ndof)).tocsr(), vals, flags.value


def mixed_mesh(n=5):
    md = meshes.plate(n)
    nid = np.arange((n + 1) ** 2).reshape(n + 1, n + 1)
    md.cnct_beams = np.concatenate([np.stack([nid[:, :-1].ravel(), nid[:, 1:].ravel()], 1),
                                    np.stack([nid[:-1, :].ravel(), nid[1:, :].ravel()], 1)]).astype(np.int32)
    md.prop_beams = np.tile([3.79e9, 3.79e9 / 2.6, 6.7e-5, 1.7e-5, 8.4e-5, 0.02], (md.cnct_beams.shape[0], 1))
    return md


@pytest.mark.parametrize('case', ['plate6', 'mixed', 'frames'])
@pytest.mark.parametrize('path', [0, 1])
def test_emulated_assembly_matches_oracle(emu2, case, path):
    """quad_geometry_kernel + assemble_tasks_kernel (path 0) and assemble_fused_kernel (path 1), run thread by
    thread on the CPU: K within 1e-10 of the oracle's COO sum, every stored value written, both paths bitwise equal."""
    md = {'plate6': lambda: meshes.plate(6), 'mixed': mixed_mesh, 'frames': lambda: meshes.frames(2, 6)}[case]()
    K, vals, flags = emu_K(emu2, md, path)
    assert not np.isnan(vals).any()
    Kref = orc.K_global(to_oracle_mesh(md))
    assert abs(K - Kref).max() / abs(Kref).max() <= 1e-10
    assert flags & 1 == 0
    if path == 0:
        _, vals1, _ = emu_K(emu2, md, 1)
        assert np.abs(vals - vals1).max() <= 1e-13 * np.abs(vals1).max()
        _, vals_b, _ = emu_K(emu2, md, 0, task_ctas=1)     # independent of the persistent grid size
        assert np.array_equal(vals, vals_b)


def test_emulated_assembly_imposes_boundary_conditions(emu2):
    md = meshes.plate(5)
    K, _, _ = emu_K(emu2, md, 0, apply_bc=True)
    Kref = orc.K_global(to_oracle_mesh(md)).toarray()
    kn = md.known
    Kref[kn, :] = 0.0; Kref[:, kn] = 0.0; Kref[kn, kn] = 1.0
    assert np.abs(K.toarray() - Kref).max() <= 1e-10 * np.abs(Kref).max()


@pytest.mark.parametrize('with_props', [True, False])
def test_emulated_adjoint_matches_complex_step(emu2, with_props):
    """quad_adjoint_kernel (cp.async pipeline as synchronous copies), beam_adjoint_kernel, node_gather_kernel."""
    md = mixed_mesh(5)
    rng = np.random.default_rng(11)
    md.crds[:, 2] += 0.05 * rng.standard_normal(md.n_node)
    u, lam = rng.standard_normal(md.ndof), rng.standard_normal(md.ndof)
    dc = np.full((md.n_node, 3), np.nan)
    dq = np.full((md.n_quad, 5), np.nan) if with_props else None
    db = np.full((md.n_beam, 6), np.nan) if with_props else None
    rc = emu2.emu_adjoint(md.n_node, md.n_quad, P(md.cnct_quads), md.n_beam, P(md.cnct_beams), P(md.crds),
                          P(md.prop_quads), P(md.prop_beams), P(u), P(lam), P(dc), P(dq), P(db), 2)
    assert rc == 0
    rdc, rdq, rdb = orc.element_sensitivity(to_oracle_mesh(md), u, lam)
    assert np.abs(dc - rdc).max() <= 1e-6 * np.abs(rdc).max()
    if with_props:
        for k in range(5):
            assert np.abs(dq[:, k] - rdq[:, k]).max() <= 1e-6 * max(np.abs(rdq[:, k]).max(), 1e-300)
        for k in range(6):
            assert np.abs(db[:, k] - rdb[:, k]).max() <= 1e-6 * max(np.abs(rdb[:, k]).max(), 1e-300)
