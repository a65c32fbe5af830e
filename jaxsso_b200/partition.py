"""Domain decomposition of one mesh over N GPUs (one process per GPU).

Nodes (= block rows of K) are owned by exactly one rank; a rank evaluates every
element that touches one of its nodes (ghost elements are recomputed, so Ke,
assembly and the adjoint reduction need no communication); columns of ghost nodes
are kept, and a 6-dof vector is completed by one halo exchange per SpMV.

Ownership comes from recursive coordinate bisection (deterministic; no METIS in the
image).  Every rank computes the same partition from the global mesh, so the halo
plans of the two sides of each interface agree without communication.
"""
from __future__ import annotations

import numpy as np

from .meshes import MeshData


def rcb_owner(crds, n_part):
    """Recursive coordinate bisection: owner rank per node (n_part a power of two)."""
    assert n_part >= 1 and (n_part & (n_part - 1)) == 0, 'n_part must be a power of two'
    owner = np.zeros(crds.shape[0], np.int32)

    def split(idx, base, parts):
        if parts == 1:
            owner[idx] = base
            return
        ext = crds[idx].max(0) - crds[idx].min(0)
        ax = int(np.argmax(ext))
        order = idx[np.argsort(crds[idx, ax], kind='stable')]
        half = order.shape[0] // 2
        split(np.sort(order[:half]), base, parts // 2)
        split(np.sort(order[half:]), base + parts // 2, parts // 2)

    split(np.arange(crds.shape[0]), 0, n_part)
    return owner


class LocalMesh:
    """What one rank holds: a MeshData in local numbering (owned nodes first, then the
    ghosts grouped by owner), plus the halo plan for jsso_set_halo."""

    def __init__(self, md, n_owned, l2g, peer_rank, send_ptr, send_idx, recv_start, recv_count,
                 quad_ids, beam_ids, remote_start):
        self.md, self.n_owned, self.l2g = md, n_owned, l2g
        self.peer_rank, self.send_ptr, self.send_idx = peer_rank, send_ptr, send_idx
        self.recv_start, self.recv_count = recv_start, recv_count
        self.quad_ids, self.beam_ids = quad_ids, beam_ids
        # where my interface nodes sit in each peer's local numbering (for direct peer-memory pushes)
        self.remote_start = remote_start


def _ghost_groups(md, owner, rank):
    """Elements touching rank's nodes, and its ghost nodes grouped by owner (ascending ids)."""
    own = owner == rank
    qsel = np.flatnonzero(own[md.cnct_quads].any(1)) if md.n_quad else np.zeros(0, np.int64)
    bsel = np.flatnonzero(own[md.cnct_beams].any(1)) if md.n_beam else np.zeros(0, np.int64)
    touched = np.unique(np.concatenate([md.cnct_quads[qsel].ravel(), md.cnct_beams[bsel].ravel()]))
    ghosts = touched[~own[touched]]
    groups = {int(p): ghosts[owner[ghosts] == p] for p in np.unique(owner[ghosts])}
    return qsel, bsel, groups


def local_mesh(md: MeshData, owner, rank, n_rank):
    owned = np.flatnonzero(owner == rank)
    qsel, bsel, groups = _ghost_groups(md, owner, rank)
    peers = sorted(groups)
    l2g = np.concatenate([owned] + [groups[p] for p in peers]).astype(np.int64)
    g2l = -np.ones(md.n_node, np.int64)
    g2l[l2g] = np.arange(l2g.shape[0])
    recv_start, recv_count, pos = [], [], owned.shape[0]
    for p in peers:
        recv_start.append(pos)
        recv_count.append(groups[p].shape[0])
        pos += groups[p].shape[0]
    # what each peer needs from me = its ghost group owned by me (same deterministic order)
    send_ptr, send_idx = [0], []
    send_peers, remote_start = [], []
    for p in range(n_rank):
        if p == rank:
            continue
        _, _, gp = _ghost_groups(md, owner, p)
        if rank in gp:
            send_peers.append(p)
            send_idx.append(g2l[gp[rank]])
            send_ptr.append(send_ptr[-1] + gp[rank].shape[0])
            remote_start.append(int(np.count_nonzero(owner == p)) +
                                sum(gp[q].shape[0] for q in gp if q < rank))
    # symmetric adjacency is guaranteed (an element touching nodes of r and p makes each a ghost of the other)
    assert send_peers == peers, (send_peers, peers)
    known_g = md.known.astype(np.int64)
    kl = g2l[known_g // 6]
    known_l = (6 * kl[kl >= 0] + (known_g % 6)[kl >= 0]).astype(np.int32)
    loads = md.loads.reshape(-1, 6)[l2g].reshape(-1)
    sub = MeshData(crds=md.crds[l2g], cnct_quads=g2l[md.cnct_quads[qsel]], prop_quads=md.prop_quads[qsel],
                   cnct_beams=g2l[md.cnct_beams[bsel]], prop_beams=md.prop_beams[bsel], known=known_l,
                   loads=loads, design_nodes=np.zeros(0, np.int64))
    return LocalMesh(sub, owned.shape[0], l2g, np.array(peers, np.int32), np.array(send_ptr, np.int32),
                     (np.concatenate(send_idx) if send_idx else np.zeros(0)).astype(np.int32),
                     np.array(recv_start, np.int32), np.array(recv_count, np.int32), qsel, bsel,
                     np.array(remote_start, np.int32))
