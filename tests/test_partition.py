"""CPU tests of the multi-GPU host logic: ownership, local meshes, halo plans.  The halo
exchange itself is replayed over torch.distributed `gloo` with world_size 2 (and 4), and the
distributed block-row SpMV is compared with the global product."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from jaxsso_b200 import meshes, partition
from oracle import jaxsso_oracle as orc
from tests.conftest import to_oracle_mesh


def test_rcb_is_balanced_and_deterministic():
    md = meshes.plate(16)
    for n in (1, 2, 4, 8):
        o = partition.rcb_owner(md.crds[:, :2], n)
        cnt = np.bincount(o, minlength=n)
        assert cnt.sum() == md.n_node and cnt.max() - cnt.min() <= n
        assert np.array_equal(o, partition.rcb_owner(md.crds[:, :2], n))


@pytest.mark.parametrize('n_rank', [2, 4])
def test_local_meshes_cover_the_mesh(n_rank):
    md = meshes.plate(12)
    owner = partition.rcb_owner(md.crds[:, :2], n_rank)
    lms = [partition.local_mesh(md, owner, r, n_rank) for r in range(n_rank)]
    owned = np.concatenate([lm.l2g[:lm.n_owned] for lm in lms])
    assert np.array_equal(np.sort(owned), np.arange(md.n_node))
    for r, lm in enumerate(lms):
        # every element touching an owned node is present, coordinates/properties follow the renumbering
        touch = (owner[md.cnct_quads] == r).any(1)
        assert np.array_equal(lm.quad_ids, np.flatnonzero(touch))
        assert np.array_equal(lm.l2g[lm.md.cnct_quads], md.cnct_quads[lm.quad_ids])
        assert np.array_equal(lm.md.crds, md.crds[lm.l2g])
        # send plan of r towards p mirrors the receive plan of p from r
        for i, p in enumerate(lm.peer_rank):
            other = lms[p]
            j = list(other.peer_rank).index(r)
            sent = lm.l2g[lm.send_idx[lm.send_ptr[i]:lm.send_ptr[i + 1]]]
            recv = other.l2g[other.recv_start[j]:other.recv_start[j] + other.recv_count[j]]
            assert np.array_equal(sent, recv)
            # direct peer-memory pushes land where the peer expects my nodes
            assert lm.remote_start[i] == other.recv_start[j]
        # prescribed dofs restricted to local nodes
        g = (6 * lm.l2g[lm.md.known // 6] + lm.md.known % 6)
        assert set(g.tolist()) <= set(md.known.tolist())


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    md = meshes.plate(10)
    K = orc.K_global(to_oracle_mesh(md)).tocsr()
    x = np.random.default_rng(5).standard_normal(md.ndof)
    owner = partition.rcb_owner(md.crds[:, :2], world)
    lm = partition.local_mesh(md, owner, rank, world)
    n_loc = lm.l2g.shape[0]
    xl = np.zeros((n_loc, 6))
    xl[:lm.n_owned] = x.reshape(-1, 6)[lm.l2g[:lm.n_owned]]        # owned part only; ghosts via exchange
    reqs, bufs = [], []
    for i, p in enumerate(lm.peer_rank):
        sb = torch.from_numpy(xl[lm.send_idx[lm.send_ptr[i]:lm.send_ptr[i + 1]]].copy())
        rb = torch.zeros((int(lm.recv_count[i]), 6), dtype=torch.float64)
        reqs.append(dist.isend(sb, int(p)))
        reqs.append(dist.irecv(rb, int(p)))
        bufs.append((i, rb, sb))
    for rq in reqs:
        rq.wait()
    for i, rb, _ in bufs:
        xl[lm.recv_start[i]:lm.recv_start[i] + lm.recv_count[i]] = rb.numpy()
    # local block rows: rows = owned global dofs, columns = local numbering
    rows = (6 * lm.l2g[:lm.n_owned, None] + np.arange(6)).ravel()
    cols = (6 * lm.l2g[:, None] + np.arange(6)).ravel()
    Kl = K[rows][:, cols]
    # no coupling outside the local column set (ghost layer is complete)
    assert abs(K[rows]).sum() == pytest.approx(abs(Kl).sum(), rel=1e-14)
    yl = Kl @ xl.ravel()
    yref = (K @ x)[rows]
    err = np.abs(yl - yref).max() / np.abs(yref).max()
    # the two scalar all-reduces of a CG step
    t = torch.tensor([float(yl @ xl[:lm.n_owned].ravel())], dtype=torch.float64)
    dist.all_reduce(t)
    ret[rank] = (err, float(t.item()), float(x @ (K @ x)))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_halo_exchange_and_distributed_spmv_gloo(world):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        err, pq, pq_ref = ret[r]
        assert err < 1e-13
        assert abs(pq - pq_ref) / abs(pq_ref) < 1e-12
