"""Vertex partitioning of a PackedGraph across the GPUs of one box + halo maps.

Each rank owns a set of vertices and therefore their half-edges (``xc``, ``mu``).  Every edge
incident to an owned vertex is local; for a cut edge the remote endpoint's half-edge is mirrored
in a *ghost* slot that the owner refreshes every iteration (5 doubles).  Both sides of a cut edge
then compute the same ``z_e = 0.5 * (xc_tail + xc_head)`` (fixed operand order: tail, head), so the
edge variables stay bitwise identical on the two ranks and only one exchange per iteration is
needed; the edge's contribution to the z-norms is accounted by the rank that owns its tail.

The reference has no distributed layer (single process, Drake thread pool);
this is new plumbing around the same iteration (``admm_solver_v3.py:655-733``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["partition_vertices", "LocalProblem", "split_graph"]


def partition_vertices(g, R, method="coord"):
    """part[v] in [0, R): contiguous strips of (almost) equal vertex count.

    ``coord``: sort by the first coordinate of each polytope's interior point (strips across the
    workspace — nearest-neighbour halos for planar problems); ``index``: split the vertex list."""
    nV = g.nV
    if R <= 1:
        return np.zeros(nV, dtype=np.int32)
    if method == "coord":
        c = g.interior_points()
        order = np.lexsort((c[:, 1], c[:, 0]))
    elif method == "index":
        order = np.arange(nV)
    else:
        raise ValueError(method)
    part = np.empty(nV, dtype=np.int32)
    bounds = np.linspace(0, nV, R + 1).astype(np.int64)
    for r in range(R):
        part[order[bounds[r]:bounds[r + 1]]] = r
    return part


class LocalProblem:
    """One rank's share of the graph, with the attribute names ``lib.graph_struct`` consumes."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def interior_points(self):
        return self.cent


def split_graph(g, part, R=None):
    """-> list of LocalProblem, one per rank."""
    part = np.asarray(part, dtype=np.int64)
    R = int(R if R is not None else part.max() + 1)
    nV, nE = g.nV, g.nE
    he_owner = g.he_owner.astype(np.int64)
    tail, head = g.edge_tail.astype(np.int64), g.edge_head.astype(np.int64)
    cent = g.interior_points()
    he_off = g.he_off.astype(np.int64)
    poly_off = g.poly_off.astype(np.int64)
    he_rank = part[he_owner]
    out = []
    for r in range(R):
        lv = np.nonzero(part == r)[0]                      # local vertices, global order
        vmap = -np.ones(nV, dtype=np.int64); vmap[lv] = np.arange(lv.shape[0])
        emask = (part[tail] == r) | (part[head] == r)
        le = np.nonzero(emask)[0]                          # local edges, global order
        emap = -np.ones(nE, dtype=np.int64); emap[le] = np.arange(le.shape[0])
        # owned half-edges: CSR of the local vertices
        deg = (he_off[lv + 1] - he_off[lv])
        l_he_off = np.zeros(lv.shape[0] + 1, dtype=np.int64); np.cumsum(deg, out=l_he_off[1:])
        own_h = np.concatenate([np.arange(he_off[v], he_off[v + 1]) for v in lv]) if lv.shape[0] else np.zeros(0, np.int64)
        nH_own = own_h.shape[0]
        hmap = -np.ones(2 * nE, dtype=np.int64); hmap[own_h] = np.arange(nH_own)
        # ghost half-edges: remote half-edges of local edges, grouped by owner rank, then by global id
        gt, gh = g.edge_he_tail.astype(np.int64)[le], g.edge_he_head.astype(np.int64)[le]
        cand = np.concatenate([gt, gh])
        ghost_h = cand[he_rank[cand] != r]
        order = np.lexsort((ghost_h, he_rank[ghost_h]))
        ghost_h = ghost_h[order]
        hmap[ghost_h] = nH_own + np.arange(ghost_h.shape[0])
        recv_counts = np.bincount(he_rank[ghost_h], minlength=R)
        # half-edges I must send: my own half-edges of cut edges, grouped by destination rank, by global id
        other_rank = np.where(part[tail[le]] == r, part[head[le]], part[tail[le]])     # remote endpoint's rank (or r)
        mine_h = np.where(part[tail[le]] == r, gt, gh)                                  # my half-edge of that edge
        cut = other_rank != r
        # an edge with both endpoints local has no remote side
        both = (part[tail[le]] == r) & (part[head[le]] == r)
        cut &= ~both
        send_h, send_to = mine_h[cut], other_rank[cut]
        order = np.lexsort((send_h, send_to))
        send_h, send_to = send_h[order], send_to[order]
        send_counts = np.bincount(send_to, minlength=R)
        lp = LocalProblem(
            rank=r, R=R, nV=int(lv.shape[0]), nE=int(le.shape[0]), nH_ghost=int(ghost_h.shape[0]),
            global_vertices=lv, global_edges=le, global_he=own_h,
            poly_off=np.concatenate([[0], np.cumsum(poly_off[lv + 1] - poly_off[lv])]).astype(np.int32),
            polyA=np.concatenate([g.polyA[poly_off[v]:poly_off[v + 1]] for v in lv]) if lv.shape[0] else np.zeros((0, 2)),
            polyb=np.concatenate([g.polyb[poly_off[v]:poly_off[v + 1]] for v in lv]) if lv.shape[0] else np.zeros(0),
            he_off=l_he_off.astype(np.int32), he_edge=emap[g.he_edge.astype(np.int64)[own_h]].astype(np.int32),
            he_flags=g.he_flags[own_h].copy(), he_out=g.he_out[own_h].copy(),
            edge_he_tail=hmap[gt].astype(np.int32), edge_he_head=hmap[gh].astype(np.int32),
            edge_counted=(part[tail[le]] == r).astype(np.uint8),
            vtype=g.vtype[lv].copy(), cent=np.ascontiguousarray(cent[lv]),
            src=int(vmap[g.src]) if g.src >= 0 else -1, dst=int(vmap[g.dst]) if g.dst >= 0 else -1,
            n_x_global=9 * nV + 18 * nE, n_mu_global=10 * nE,
            send_idx=hmap[send_h].astype(np.int64), send_counts=send_counts.astype(np.int64),
            recv_counts=recv_counts.astype(np.int64),
            max_live_degree=g.max_live_degree, max_rows=g.max_rows,
        )
        assert np.all(lp.edge_he_tail >= 0) and np.all(lp.edge_he_head >= 0)
        out.append(lp)
    # consistency: what r sends to q is what q expects from r
    for r in range(R):
        for q in range(R):
            assert out[r].send_counts[q] == out[q].recv_counts[r]
    return out
