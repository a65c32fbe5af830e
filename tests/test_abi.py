"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol
include/jsso.h declares, the host symbolic pass is bit-exact against the oracle's
sorted-unique COO pattern, and compute entry points refuse to run without a GPU
(no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from jaxsso_b200 import _native as nat
from jaxsso_b200 import build as jbuild
from jaxsso_b200 import meshes
from oracle import jaxsso_oracle as orc
from tests.conftest import ROOT, to_oracle_mesh


@pytest.fixture(scope='module', autouse=True)
def built():
    jbuild.build()


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'jsso.h')).read()
    declared = set(re.findall(r'\b(jsso_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)
    L = nat.lib()
    for s in declared:
        assert hasattr(L, s), s


def symbolic_handle(md, n_row=None):
    return nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=-1, n_row=n_row)


def expand(rowptr, colidx):
    """6x6 block pattern -> lexicographically sorted scalar (row, col) pairs."""
    nb = np.diff(rowptr)
    brow = np.repeat(np.arange(rowptr.shape[0] - 1), nb)
    rows, cols = [], []
    r = (6 * brow[:, None, None] + np.arange(6)[None, :, None] + np.zeros((1, 1, 6), int)).reshape(-1)
    c = (6 * colidx.astype(np.int64)[:, None, None] + np.zeros((1, 6, 1), int) + np.arange(6)[None, None, :]).reshape(-1)
    key = np.sort(r.astype(np.int64) * (6 * (rowptr.shape[0] - 1)) + c)
    return key


@pytest.mark.parametrize('name', ['barrel_arch', 'beam_arch', 'frames10', 'plate8', 'mannheim', 'mixed'])
def test_pattern_bit_exact(name, mannheim_data):
    if name == 'barrel_arch':
        md = meshes.barrel_arch()
    elif name == 'beam_arch':
        md = meshes.beam_arch()
    elif name == 'frames10':
        md = meshes.frames(10, 100)
    elif name == 'plate8':
        md = meshes.plate(8)
    elif name == 'mannheim':
        md = meshes.mannheim_quad(mannheim_data)
    else:
        md = meshes.plate(6)
        nid = np.arange(49).reshape(7, 7)
        md.cnct_beams = np.stack([nid[:, :-1].ravel(), nid[:, 1:].ravel()], 1).astype(np.int32)
        md.prop_beams = np.tile([1e9, 4e8, 1e-5, 2e-5, 3e-5, 1e-2], (md.cnct_beams.shape[0], 1))
    h = symbolic_handle(md)
    rowptr, colidx = h.pattern()
    assert h.n_items == 16 * md.n_quad + 4 * md.n_beam
    # sorted within rows, no duplicates
    for r in range(h.n_row):
        c = colidx[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(c) > 0)
    pr, pc = orc.sorted_unique_pattern(to_oracle_mesh(md))
    assert np.array_equal(expand(rowptr, colidx), pr * md.ndof + pc)
    h.close()


def test_symbolic_rejects_bad_mesh():
    with pytest.raises(nat.JssoError):
        nat.Handle(4, np.array([[0, 1, 2, 7]]), None, None, device=-1)
    with pytest.raises(nat.JssoError):
        nat.Handle(4, np.array([[0, 1, 2, 3]]), None, np.array([24]), device=-1)


def test_no_cpu_fallback():
    md = meshes.plate(4)
    h = symbolic_handle(md)
    with pytest.raises(nat.JssoError) as ei:
        h.assemble(None, None, None)
    assert ei.value.code == 9
    if nat.lib().jsso_device_count() == 0:
        with pytest.raises(nat.JssoError) as ei:
            nat.Handle(md.n_node, md.cnct_quads, None, md.known, device=0)
        assert ei.value.code == 2 and 'no CPU fallback' in str(ei.value)


def test_owned_rows_subset():
    """n_row < n_node (multi-GPU local mesh): rows only for owned nodes, columns may be ghosts."""
    md = meshes.plate(6)
    full = symbolic_handle(md)
    part = symbolic_handle(md, n_row=20)
    rp, ci = full.pattern()
    rp2, ci2 = part.pattern()
    assert np.array_equal(rp2, rp[:21]) and np.array_equal(ci2, ci[:rp[20]])
