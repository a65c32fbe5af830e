"""CPU checks of the warp-task lists of the two-kernel assembly (symbolic pass, no GPU):
the task invariants the CUDA kernel relies on, and a NumPy replay of the kernel's indexing
(item -> lane, first-item ballot, per-block sums in list order, boundary word) fed with the
ORACLE's element matrices, compared with the oracle's assembled K."""
import numpy as np
import pytest

from jaxsso_b200 import _native as nat
from jaxsso_b200 import build as jbuild
from jaxsso_b200 import meshes
from oracle import jaxsso_oracle as orc
from tests.conftest import to_oracle_mesh


@pytest.fixture(scope='module', autouse=True)
def built():
    jbuild.build()


def mixed_mesh():
    md = meshes.plate(6)
    nid = np.arange(49).reshape(7, 7)
    md.cnct_beams = np.stack([nid[:, :-1].ravel(), nid[:, 1:].ravel()], 1).astype(np.int32)
    md.prop_beams = np.tile([1e9, 4e8, 1e-5, 2e-5, 3e-5, 1e-2], (md.cnct_beams.shape[0], 1))
    return md


def make(name, mannheim_data):
    return {'barrel_arch': meshes.barrel_arch, 'beam_arch': meshes.beam_arch,
            'frames10': lambda: meshes.frames(10, 100), 'plate8': lambda: meshes.plate(8),
            'mannheim': lambda: meshes.mannheim_quad(mannheim_data), 'mixed': mixed_mesh}[name]()


def nth_set_bit(mask, n):
    """Position of the n-th (1-based) set bit of mask, as CUDA's __fns(mask, 0, n)."""
    for pos in range(32):
        if (mask >> pos) & 1:
            n -= 1
            if n == 0:
                return pos
    return 0xffffffff


def replay(md, t, ke_q, ke_b, apply_bc):
    """What assemble_tasks_kernel computes, lane by lane: (nnzb, 6, 6) blocks in [row, col] orientation."""
    nq = md.n_quad
    vals = np.full((t['blk_bc'].shape[0], 36), np.nan)
    for blk0, item0, el0, packed in t['task_meta']:
        n_blk, n_item, n_el = packed & 255, (packed >> 8) & 255, (packed >> 16) & 255
        desc = t['item_desc'][item0:item0 + n_item].astype(np.int64)
        buf = np.zeros((n_item, 36))
        firsts = 0
        for lane in range(n_item):
            d = int(desc[lane])
            lb, b, a, lq = d & 31, (d >> 5) & 3, (d >> 7) & 3, (d >> 9) & 7
            el = t['item_code'][item0 + lane] >> 4
            if d & (1 << 12):
                blk = ke_b[el - nq][6 * a:6 * a + 6, 6 * b:6 * b + 6]
            else:
                assert lq < n_el and t['task_els'][el0 + lq] == el
                blk = ke_q[el][6 * a:6 * a + 6, 6 * b:6 * b + 6]
            buf[lane] = blk.T.reshape(-1)            # column-major: out[6 j + i]
            if d & (1 << 13):
                firsts |= 1 << lane
        st = [nth_set_bit(firsts, l + 1) if l < n_blk else n_item for l in range(33)]
        for bl in range(n_blk):
            s0, s1 = st[bl], (st[bl + 1] if bl + 1 < n_blk else n_item)
            v = buf[s0].copy()
            for it in range(s0 + 1, s1):
                v += buf[it]
            bc = int(t['blk_bc'][blk0 + bl]) if apply_bc else 0
            if bc & 0xfff:
                rm, cm, diag = bc & 63, (bc >> 6) & 63, (bc >> 12) & 1
                for j in range(6):
                    for i in range(6):
                        if (rm >> i) & 1 or (cm >> j) & 1:
                            v[6 * j + i] = 1.0 if (diag and i == j) else 0.0
            vals[blk0 + bl] = v
    assert not np.isnan(vals).any()                  # every block written exactly by one task
    return vals.reshape(-1, 6, 6).transpose(0, 2, 1)


@pytest.mark.parametrize('name', ['barrel_arch', 'beam_arch', 'frames10', 'plate8', 'mannheim', 'mixed'])
def test_task_invariants_and_replay(name, mannheim_data):
    md = make(name, mannheim_data)
    h = nat.Handle(md.n_node, md.cnct_quads, md.cnct_beams, md.known, device=-1)
    t = h.assembly_tasks()
    assert t['tasks_ok']
    meta = t['task_meta']
    n_blk, n_item, n_el = meta[:, 3] & 255, (meta[:, 3] >> 8) & 255, (meta[:, 3] >> 16) & 255
    assert (n_item >= 1).all() and (n_item <= 32).all() and (n_el <= 8).all() and (n_blk >= 1).all() and (n_blk <= 32).all()
    # tasks tile the block and item ranges without gaps, in order
    assert meta[0, 0] == 0 and meta[0, 1] == 0 and meta[0, 2] == 0
    assert np.array_equal(meta[1:, 0], meta[:-1, 0] + n_blk[:-1]) and meta[-1, 0] + n_blk[-1] == h.nnzb
    assert np.array_equal(meta[1:, 1], meta[:-1, 1] + n_item[:-1]) and meta[-1, 1] + n_item[-1] == h.n_items
    assert np.array_equal(meta[1:, 2], meta[:-1, 2] + n_el[:-1]) and meta[-1, 2] + n_el[-1] == t['task_els'].shape[0]
    assert np.array_equal(t['blk_item_ptr'][meta[:, 0]], meta[:, 1])
    # first-item flags = starts of the contributor lists; local block ids count up from 0 in every task
    first = (t['item_desc'] >> 13) & 1
    assert np.array_equal(np.flatnonzero(first), t['blk_item_ptr'][:-1])
    # oracle element matrices -> replay -> compare with the oracle's K
    m = to_oracle_mesh(md)
    ke_q = orc.element_K_quad(md.crds[md.cnct_quads].reshape(-1, 12), md.prop_quads) if md.n_quad else np.zeros((0, 24, 24))
    ke_b = orc.element_K_beamcol(md.crds[md.cnct_beams].reshape(-1, 6), md.prop_beams) if md.n_beam else np.zeros((0, 12, 12))
    rp, ci = h.pattern()
    K = orc.K_global(m).tocsr()
    blocks = replay(md, t, ke_q, ke_b, apply_bc=False)
    Kt = nat.bsr_to_scipy(rp, ci, blocks).tocsr()
    assert abs(Kt - K).max() <= 1e-12 * abs(K).max()
    # boundary conditions: prescribed rows/cols -> identity
    blocks_bc = replay(md, t, ke_q, ke_b, apply_bc=True)
    Kb = nat.bsr_to_scipy(rp, ci, blocks_bc).tolil()
    Kr = K.tolil()
    known = np.asarray(md.known, int)
    Kr[known, :] = 0.0
    Kr[:, known] = 0.0
    Kr[known, known] = 1.0
    assert abs(Kb.tocsr() - Kr.tocsr()).max() <= 1e-12 * abs(K).max()
    h.close()


def test_tasks_fall_back_on_empty_rows():
    """A node without elements has an empty diagonal block: the chunked kernel handles that mesh."""
    md = meshes.plate(4)
    h = nat.Handle(md.n_node + 1, md.cnct_quads, md.cnct_beams, md.known, device=-1)
    t = h.assembly_tasks()
    assert not t['tasks_ok'] and t['n_task'] == 0
    h.close()
