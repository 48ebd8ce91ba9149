"""Host-side mesh partition for multi-GPU runs.

Vertices (the owners of the (d+1)-unknown blocks) are split by recursive coordinate bisection; every rank
gets its owned vertices, *all* cells touching one of them (one-layer overlap, so owned rows assemble
locally with no assembly communication) and the ghost vertices those cells bring in, grouped by owner.
The only data-path communication this implies is the ghost-value halo exchange before an operator
application and the allreduce of Krylov dot products.

The reference delegates partitioning to DOLFIN/SCOTCH; that partition is not reproducible without SCOTCH
(SURVEY.md 8c.3), so this is *a* deterministic partition, not the reference's.
"""
import numpy as np


def rcb(coords, n_parts):
    """Recursive coordinate bisection into n_parts (any integer >= 1) of near-equal size. Deterministic."""
    part = np.zeros(len(coords), dtype=np.int32)

    def split(idx, p0, k):
        if k == 1:
            part[idx] = p0
            return
        k0 = k // 2
        X = coords[idx]
        axis = int(np.argmax(X.max(axis=0) - X.min(axis=0)))
        n0 = int(round(len(idx) * k0 / k))
        order = np.argsort(X[:, axis], kind="stable")
        split(idx[order[:n0]], p0, k0)
        split(idx[order[n0:]], p0 + k0, k - k0)

    split(np.arange(len(coords)), 0, int(n_parts))
    return part


class LocalMesh:
    """One rank's share. Local vertex numbering: [owned (ascending global id) | ghosts grouped by owner rank]."""

    def __init__(self, rank, n_ranks, dim, l2g, n_owned, cells_local, cell_ids, peers, send_ptr, send_idx, recv_ptr):
        self.rank, self.n_ranks, self.dim = rank, n_ranks, dim
        self.l2g, self.n_owned = l2g, n_owned
        self.cells, self.cell_ids = cells_local, cell_ids
        self.peers, self.send_ptr, self.send_idx, self.recv_ptr = peers, send_ptr, send_idx, recv_ptr

    @property
    def n_local(self):
        return len(self.l2g)

    def to_local(self, x_global, nb=None):
        nb = self.dim + 1 if nb is None else nb
        return np.ascontiguousarray(np.asarray(x_global).reshape(-1, nb)[self.l2g].ravel())

    def local_dofs(self, global_dofs, vals, nb=None):
        """Global (dof, value) pairs restricted to local vertices (owned + ghost), in local numbering."""
        nb = self.dim + 1 if nb is None else nb
        g2l = self.global_to_local()
        v, k = np.asarray(global_dofs) // nb, np.asarray(global_dofs) % nb
        loc = g2l[v]
        keep = loc >= 0
        return (loc[keep] * nb + k[keep]).astype(np.int64), np.asarray(vals)[keep]

    def global_to_local(self):
        if not hasattr(self, "_g2l"):
            n_global = int(self._n_global)
            g2l = np.full(n_global, -1, dtype=np.int64)
            g2l[self.l2g] = np.arange(len(self.l2g))
            self._g2l = g2l
        return self._g2l


def build_local_mesh(coords, cells, part, rank, n_ranks):
    """Sub-mesh + halo plan of ``rank`` from the global mesh and the vertex->rank map (every rank can call this
    on the full mesh; only partition-boundary cells are inspected for the plan)."""
    cells = np.asarray(cells)
    nv = cells.shape[1]
    pc = part[cells]                                         # owner of each cell vertex
    mine = (pc == rank).any(axis=1)
    cell_ids = np.nonzero(mine)[0]
    my_cells = cells[cell_ids]
    owned = np.nonzero(part == rank)[0]
    used = np.unique(my_cells)
    ghosts = used[part[used] != rank]
    gorder = np.lexsort((ghosts, part[ghosts]))              # by owner rank, then global id
    ghosts = ghosts[gorder]
    l2g = np.concatenate([owned, ghosts]).astype(np.int64)
    g2l = np.full(len(part), -1, dtype=np.int64)
    g2l[l2g] = np.arange(len(l2g))
    cells_local = g2l[my_cells].astype(np.int32)
    # receive plan: ghosts are contiguous per owner
    gowner = part[ghosts]
    recv_peers = np.unique(gowner)
    # send plan: my owned vertices that appear in a cell together with a vertex owned by q are ghosts of q
    mixed = (pc != pc[:, :1]).any(axis=1) & mine
    mc, mp = cells[mixed], pc[mixed]
    need = []
    for a in range(nv):
        for b in range(nv):
            if a == b:
                continue
            sel = (mp[:, a] != rank) & (mp[:, b] == rank)
            need.append(np.stack([mp[sel, a].astype(np.int64), mc[sel, b].astype(np.int64)], axis=1))
    need = np.unique(np.concatenate(need, axis=0), axis=0) if need else np.zeros((0, 2), dtype=np.int64)
    send_peers = np.unique(need[:, 0]).astype(np.int64)
    peers = np.union1d(recv_peers, send_peers).astype(np.int32)
    send_ptr, recv_ptr, send_idx = [0], [0], []
    for q in peers:
        s = need[need[:, 0] == q, 1]                         # ascending global id == q's ghost order for my vertices
        send_idx.append(g2l[s])
        send_ptr.append(send_ptr[-1] + len(s))
        recv_ptr.append(recv_ptr[-1] + int((gowner == q).sum()))
    send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, dtype=np.int32)
    lm = LocalMesh(rank, n_ranks, coords.shape[1], l2g, len(owned), cells_local, cell_ids, peers,
                   np.asarray(send_ptr, dtype=np.int64), send_idx, np.asarray(recv_ptr, dtype=np.int64))
    lm._n_global = len(part)
    lm.coords = np.ascontiguousarray(coords[l2g])
    return lm
