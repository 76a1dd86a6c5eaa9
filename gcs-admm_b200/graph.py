"""Host-side GCS graph layer (Drake-free).

Restates the *behaviour* of the reference's graph construction
(reference ``utils.py:31-82`` ``build_graph``, ``utils.py:12-28``
``convert_pt_to_polytope``, ``utils.py:85-98`` ``delta``) without pydrake:

* vertex list  = ``list(As.keys())`` (reference ``utils.py:46``)
* directed edge ``(v1, v2)`` for every ordered pair of distinct vertices whose
  polytopes intersect, enumerated ``for v1 in V: for v2 in V`` (``utils.py:68-72``)
  so every overlap contributes two directed edges, in lexicographic
  (index(v1), index(v2)) order
* ``I_v_out[v]`` / ``I_v_in[v]`` filled in edge order (``utils.py:75-80``)

The reference decides overlap by an LP feasibility solve per ordered pair
(|V|^2 Drake solves).  Here the 2-D case is decided exactly with a
separating-axis test on the polygons' vertices behind a uniform-grid broad
phase, so 10^6 regions build in seconds; for n != 2 an LP (scipy HiGHS) is used.

The second half of the module flattens the graph into the half-edge CSR layout
the CUDA library consumes (``include/gcsadmm.h`` ``GcsGraph``).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "convert_pt_to_polytope", "delta", "polygon_vertices", "build_graph",
    "pack_polytopes", "build_graph_packed", "PackedGraph", "pack_graph", "pack_batch",
]

OVERLAP_TOL = 1e-9


def convert_pt_to_polytope(pt, eps=1e-6):
    """Axis-aligned box of half-width ``eps`` around ``pt`` as (A, b) with A x <= b.

    Same contract as reference ``utils.py:12-28``: A = [I; -I], b = [pt+eps; -pt+eps].
    """
    pt = np.asarray(pt, dtype=float)
    n = pt.shape[0]
    A = np.concatenate([np.eye(n), -np.eye(n)], axis=0)
    b = np.concatenate([pt + eps, eps - pt])
    return A, b


def delta(v1, v2):
    """delta_{v1 v2} of the GCS formulation (reference ``utils.py:85-98``):
    1 only for ('s','s') and ('t','t')."""
    if v1 == v2 and (v1 == 's' or v1 == 't'):
        return 1
    return 0


# --------------------------------------------------------------------------
# polygons
# --------------------------------------------------------------------------

def polygon_vertices(A, b, tol=1e-9):
    """Vertices of the bounded 2-D polygon {x : A x <= b}, counter-clockwise.

    All pairwise intersections of the boundary lines are formed and the
    feasible ones kept; duplicates (within ``tol``) are merged.
    """
    A = np.asarray(A, dtype=float)
    b = np.asarray(b, dtype=float)
    m = A.shape[0]
    ii, jj = np.triu_indices(m, 1)
    a1, a2 = A[ii], A[jj]
    det = a1[:, 0] * a2[:, 1] - a1[:, 1] * a2[:, 0]
    ok = np.abs(det) > 1e-14
    ii, jj, a1, a2, det = ii[ok], jj[ok], a1[ok], a2[ok], det[ok]
    px = (b[ii] * a2[:, 1] - a1[:, 1] * b[jj]) / det
    py = (a1[:, 0] * b[jj] - b[ii] * a2[:, 0]) / det
    P = np.stack([px, py], axis=1)
    scale = max(1.0, float(np.max(np.abs(b))))
    feas = np.all(P @ A.T <= b[None, :] + tol * scale, axis=1)
    P = P[feas]
    if P.shape[0] == 0:
        return P
    # merge duplicates
    keep = []
    for p in P:
        if not any(np.max(np.abs(p - q)) <= 10 * tol * scale for q in keep):
            keep.append(p)
    P = np.array(keep)
    c = P.mean(axis=0)
    ang = np.arctan2(P[:, 1] - c[1], P[:, 0] - c[0])
    return P[np.argsort(ang)]


def pack_polytopes(As, bs, keys=None):
    """Flatten dict-of-polytopes into (poly_off[nV+1], A[sum m, n], b[sum m])."""
    if keys is None:
        keys = list(As.keys())
    ms = np.array([np.asarray(As[k]).shape[0] for k in keys], dtype=np.int64)
    off = np.zeros(len(keys) + 1, dtype=np.int64)
    np.cumsum(ms, out=off[1:])
    n = np.asarray(As[keys[0]]).shape[1]
    A = np.empty((int(off[-1]), n), dtype=np.float64)
    b = np.empty(int(off[-1]), dtype=np.float64)
    for i, k in enumerate(keys):
        A[off[i]:off[i + 1]] = np.asarray(As[k], dtype=np.float64)
        b[off[i]:off[i + 1]] = np.asarray(bs[k], dtype=np.float64).reshape(-1)
    return off, A, b


def _vertices_batch(off, A, b, tol=1e-9):
    """Polygon vertices of every polytope, padded: (nV, kmax, 2) + count (nV,).

    Polytopes are grouped by row count so each group is one vectorised pass.
    Padding repeats the first vertex (harmless for min/max and SAT tests).
    """
    nV = off.shape[0] - 1
    ms = np.diff(off)
    groups = []
    cnt = np.zeros(nV, dtype=np.int64)
    for m in np.unique(ms):
        idx = np.nonzero(ms == m)[0]
        rows = off[idx][:, None] + np.arange(m)[None, :]
        Ag, bg = A[rows], b[rows]                      # (g, m, 2), (g, m)
        ii, jj = np.triu_indices(int(m), 1)
        a1, a2 = Ag[:, ii], Ag[:, jj]                  # (g, p, 2)
        det = a1[..., 0] * a2[..., 1] - a1[..., 1] * a2[..., 0]
        good = np.abs(det) > 1e-14
        dets = np.where(good, det, 1.0)
        px = (bg[:, ii] * a2[..., 1] - a1[..., 1] * bg[:, jj]) / dets
        py = (a1[..., 0] * bg[:, jj] - bg[:, ii] * a2[..., 0]) / dets
        P = np.stack([px, py], axis=-1)                # (g, p, 2)
        scale = np.maximum(1.0, np.max(np.abs(bg), axis=1))[:, None, None]
        viol = np.einsum('gpk,gmk->gpm', P, Ag) - bg[:, None, :]
        feas = good & np.all(viol <= tol * scale, axis=2)
        order = np.argsort(~feas, axis=1, kind="stable")          # feasible intersections first, original order kept
        groups.append((idx, np.take_along_axis(P, order[..., None], axis=1)))
        cnt[idx] = feas.sum(axis=1)
    kmax = max(1, int(cnt.max()) if nV else 1)
    out = np.zeros((nV, kmax, 2))
    for idx, Ps in groups:
        k = min(kmax, Ps.shape[1])
        out[idx, :k] = Ps[:, :k]
    pad = np.arange(kmax)[None, :] >= cnt[:, None]
    out = np.where(pad[..., None], out[:, :1], out)
    out[cnt == 0] = 0.0
    return out, cnt


def _sat_overlap(off, A, b, verts, pi, pj, tol):
    """Separating-axis test for polygon pairs (pi[k], pj[k]); True = they intersect.

    Two convex polygons are disjoint iff some boundary row of one has all the
    other's vertices strictly outside.  Closed sets: touching counts as overlap,
    as it does for the reference's LP feasibility test (``utils.py:49-65``).
    """
    ms = np.diff(off)
    mmax = int(ms.max())
    nV = ms.shape[0]
    # padded rows (pad rows are 0 x <= +inf -> never separating)
    Ap = np.zeros((nV, mmax, 2))
    bp = np.full((nV, mmax), np.inf)
    for m in np.unique(ms):
        idx = np.nonzero(ms == m)[0]
        rows = off[idx][:, None] + np.arange(m)[None, :]
        Ap[idx, :m] = A[rows]
        bp[idx, :m] = b[rows]
    res = np.ones(pi.shape[0], dtype=bool)
    CH = 200000
    for s in range(0, pi.shape[0], CH):
        a, c = pi[s:s + CH], pj[s:s + CH]
        for p, q in ((a, c), (c, a)):
            # rows of p against vertices of q
            val = np.einsum('kmd,kvd->kmv', Ap[p], verts[q]) - bp[p][:, :, None]
            sep = np.any(np.min(val, axis=2) > tol, axis=1)
            res[s:s + CH] &= ~sep
    return res


def build_graph_packed(off, A, b, tol=OVERLAP_TOL):
    """Directed overlap graph of packed 2-D polytopes.

    Returns ``(edge_tail, edge_head)`` int64 arrays in the reference's edge order
    (lexicographic in (tail index, head index)); see module docstring.
    """
    nV = off.shape[0] - 1
    if A.shape[1] != 2:
        return _build_graph_lp(off, A, b)
    verts, cnt = _vertices_batch(off, A, b)
    if np.any(cnt == 0):
        bad = int(np.nonzero(cnt == 0)[0][0])
        raise ValueError(f"polytope #{bad} is empty or unbounded")
    lo = verts.min(axis=1)
    hi = verts.max(axis=1)
    # broad phase: uniform grid over AABB centres, cell >= max extent
    ext = hi - lo
    cell = max(float(np.max(ext)), 1e-12)
    gmin = lo.min(axis=0)
    c0 = np.floor((lo - gmin) / cell).astype(np.int64)
    c1 = np.floor((hi - gmin) / cell).astype(np.int64)
    # every box spans at most 2 cells per axis -> register it in each spanned cell
    ent_v, ent_c = [], []
    W = int(c1[:, 0].max()) + 2
    for dx in (0, 1):
        for dy in (0, 1):
            cx = np.minimum(c0[:, 0] + dx, c1[:, 0])
            cy = np.minimum(c0[:, 1] + dy, c1[:, 1])
            use = np.ones(nV, dtype=bool)
            if dx:
                use &= c1[:, 0] > c0[:, 0]
            if dy:
                use &= c1[:, 1] > c0[:, 1]
            ent_v.append(np.nonzero(use)[0])
            ent_c.append((cy * W + cx)[use])
    ent_v = np.concatenate(ent_v)
    ent_c = np.concatenate(ent_c)
    order = np.argsort(ent_c, kind='stable')
    ent_v, ent_c = ent_v[order], ent_c[order]
    starts = np.nonzero(np.r_[True, ent_c[1:] != ent_c[:-1]])[0]
    ends = np.r_[starts[1:], ent_c.shape[0]]
    sizes = ends - starts
    # candidate pairs inside each cell, grouped by cell population
    pis, pjs = [], []
    for sz in np.unique(sizes):
        if sz < 2:
            continue
        cells = starts[sizes == sz]
        members = ent_v[cells[:, None] + np.arange(sz)[None, :]]     # (c, sz)
        ii, jj = np.triu_indices(int(sz), 1)
        pis.append(members[:, ii].ravel())
        pjs.append(members[:, jj].ravel())
    if pis:
        pi = np.concatenate(pis)
        pj = np.concatenate(pjs)
        a = np.minimum(pi, pj)
        c = np.maximum(pi, pj)
        key = np.unique(a * nV + c)
        pi, pj = key // nV, key % nV
        # AABB reject
        ok = np.all(lo[pi] <= hi[pj] + tol, axis=1) & np.all(lo[pj] <= hi[pi] + tol, axis=1)
        pi, pj = pi[ok], pj[ok]
        hit = _sat_overlap(off, A, b, verts, pi, pj, tol)
        pi, pj = pi[hit], pj[hit]
    else:
        pi = pj = np.zeros(0, dtype=np.int64)
    tail = np.concatenate([pi, pj])
    head = np.concatenate([pj, pi])
    order = np.lexsort((head, tail))
    return tail[order].astype(np.int64), head[order].astype(np.int64)


def _build_graph_lp(off, A, b):
    """General-n fallback: LP feasibility per unordered pair (scipy HiGHS)."""
    from scipy.optimize import linprog
    nV = off.shape[0] - 1
    n = A.shape[1]
    tails, heads = [], []
    for i in range(nV):
        for j in range(i + 1, nV):
            Ac = np.vstack([A[off[i]:off[i + 1]], A[off[j]:off[j + 1]]])
            bc = np.concatenate([b[off[i]:off[i + 1]], b[off[j]:off[j + 1]]])
            r = linprog(np.zeros(n), A_ub=Ac, b_ub=bc, bounds=[(None, None)] * n, method='highs')
            if r.status == 0:
                tails += [i, j]
                heads += [j, i]
    tail = np.array(tails, dtype=np.int64)
    head = np.array(heads, dtype=np.int64)
    order = np.lexsort((head, tail))
    return tail[order], head[order]


def build_graph(As, bs):
    """Drop-in for reference ``utils.py:31-82``: returns (V, E, I_v_in, I_v_out)
    with the reference's orderings (see module docstring)."""
    V = list(As.keys())
    off, A, b = pack_polytopes(As, bs, V)
    tail, head = build_graph_packed(off, A, b)
    E = [(V[i], V[j]) for i, j in zip(tail.tolist(), head.tolist())]
    I_v_in = {v: [] for v in V}
    I_v_out = {v: [] for v in V}
    for e in E:
        I_v_out[e[0]].append(e)
        I_v_in[e[1]].append(e)
    return V, E, I_v_in, I_v_out


# --------------------------------------------------------------------------
# flat half-edge layout for the device
# --------------------------------------------------------------------------

class PackedGraph:
    """Flat arrays the C-ABI consumes (``include/gcsadmm.h`` ``GcsGraph``).

    Half-edges of vertex v are enumerated in the reference's per-vertex variable
    order ``I_v_in[v] + I_v_out[v]`` (reference ``admm_solver_v3.py:105-116``).

    Attributes
    ----------
    nV, nE, n : sizes
    poly_off[nV+1], polyA[sum m * 2], polyb[sum m] : packed polytopes, V order
    he_off[nV+1] : CSR over half-edges
    he_edge[H]   : edge id of each half-edge
    he_out[H]    : 1 if the owner is the edge's tail (outgoing), 0 if head
    edge_tail[nE], edge_head[nE] : endpoints (vertex ids)
    edge_he_tail[nE], edge_he_head[nE] : half-edge id of the edge at tail / head
    src, dst : vertex ids of 's' and 't' (-1 if absent)
    """

    def __init__(self, off, A, b, tail, head, src, dst, keys=None):
        nV = off.shape[0] - 1
        nE = tail.shape[0]
        self.nV, self.nE, self.n = nV, nE, A.shape[1]
        self.keys = keys
        self.poly_off = off.astype(np.int32)
        self.polyA = np.ascontiguousarray(A, dtype=np.float64)
        self.polyb = np.ascontiguousarray(b, dtype=np.float64)
        self.edge_tail = tail.astype(np.int32)
        self.edge_head = head.astype(np.int32)
        # one (src, dst) pair for a single graph, arrays for block-diagonally packed independent problems
        self.srcs = np.atleast_1d(np.asarray(src, dtype=np.int64))
        self.dsts = np.atleast_1d(np.asarray(dst, dtype=np.int64))
        self.src, self.dst = int(self.srcs[0]), int(self.dsts[0])
        # per vertex: incoming edges (in edge order) then outgoing edges
        eid = np.arange(nE, dtype=np.int64)
        owner = np.concatenate([head, tail])
        is_out = np.concatenate([np.zeros(nE, np.int64), np.ones(nE, np.int64)])
        edges = np.concatenate([eid, eid])
        order = np.lexsort((edges, is_out, owner))
        owner, is_out, edges = owner[order], is_out[order], edges[order]
        deg = np.bincount(owner, minlength=nV)
        he_off = np.zeros(nV + 1, dtype=np.int64)
        np.cumsum(deg, out=he_off[1:])
        self.he_off = he_off.astype(np.int32)
        self.he_edge = edges.astype(np.int32)
        self.he_out = is_out.astype(np.int32)
        self.he_owner = owner.astype(np.int32)
        H = 2 * nE
        hid = np.arange(H, dtype=np.int64)
        self.edge_he_tail = np.empty(nE, dtype=np.int32)
        self.edge_he_head = np.empty(nE, dtype=np.int32)
        self.edge_he_tail[edges[is_out == 1]] = hid[is_out == 1]
        self.edge_he_head[edges[is_out == 0]] = hid[is_out == 0]
        self.d_in = np.bincount(head, minlength=nV).astype(np.int32)
        self.d_out = np.bincount(tail, minlength=nV).astype(np.int32)
        self._classify()

    def _classify(self):
        """Presolve flags the kernels consume (same rule as the oracle's ``classify``).

        he_flags bit0 = outgoing, bit1 = flow forced to 0 (in-edges of 's', out-edges of
        't', every edge of a vertex left without a live in- or out-edge).
        vtype: 0 generic, 1 source, 2 target, 3 dead (no flow can pass)."""
        nV, H = self.nV, 2 * self.nE
        owner, out = self.he_owner.astype(np.int64), self.he_out.astype(bool)
        zero = np.zeros(H, dtype=bool)
        is_src = np.zeros(nV, dtype=bool); is_src[self.srcs[self.srcs >= 0]] = True
        is_dst = np.zeros(nV, dtype=bool); is_dst[self.dsts[self.dsts >= 0]] = True
        zero |= is_src[owner] & ~out
        zero |= is_dst[owner] & out
        live_in = np.bincount(owner[~out & ~zero], minlength=nV)
        live_out = np.bincount(owner[out & ~zero], minlength=nV)
        dead = (~is_src & (live_in == 0)) | (~is_dst & (live_out == 0))
        zero |= dead[owner]
        vtype = np.zeros(nV, dtype=np.uint8)
        vtype[is_src] = 1
        vtype[is_dst] = 2
        vtype[dead] = 3
        if np.any(dead & is_src):
            raise ValueError("infeasible problem: source has no outgoing edge")
        if np.any(dead & is_dst):
            raise ValueError("infeasible problem: target has no incoming edge")
        self.he_flags = (out.astype(np.uint8) | (zero.astype(np.uint8) << 1)).astype(np.uint8)
        self.vtype = vtype
        live_deg = np.bincount(owner[~zero], minlength=nV)
        self.max_live_degree = int(live_deg.max()) if nV else 0
        self.max_rows = int(np.diff(self.poly_off).max()) if nV else 0

    @property
    def H(self):
        return 2 * self.nE

    def polygon_vertices_batch(self):
        """(verts[nV, kmax, 2], count[nV]) of every region's polygon, computed once per graph (interior points, cone tables)."""
        if getattr(self, "_verts", None) is None:
            self._verts = _vertices_batch(self.poly_off.astype(np.int64), self.polyA, self.polyb)
        return self._verts

    def interior_points(self):
        """One strictly interior point per polytope (vertex centroid); the device
        IPM starts from it and reports it as x_v of flow-less vertices.  Computed once per graph."""
        if getattr(self, "_cent", None) is not None:
            return self._cent
        verts, cnt = self.polygon_vertices_batch()
        mask = np.arange(verts.shape[1])[None, :] < cnt[:, None]
        c = (verts * mask[:, :, None]).sum(axis=1) / cnt[:, None]
        self._cent = np.ascontiguousarray(c)
        return self._cent


def pack_graph(As, bs, V=None, E=None):
    """(As, bs[, V, E]) -> PackedGraph.  Builds the graph if V/E are not given."""
    if V is None or E is None:
        V, E, _, _ = build_graph(As, bs)
    off, A, b = pack_polytopes(As, bs, V)
    if A.shape[1] != 2:
        raise ValueError("the CUDA path is specialised to n = 2 (all reference data is 2-D)")
    index = {v: i for i, v in enumerate(V)}
    tail = np.array([index[e[0]] for e in E], dtype=np.int64)
    head = np.array([index[e[1]] for e in E], dtype=np.int64)
    return PackedGraph(off, A, b, tail, head, index.get('s', -1), index.get('t', -1), keys=V)


def pack_batch(graphs):
    """Block-diagonal packing of independent problems (BASELINE config "batch of 4096 start/goal queries").

    ``graphs``: list of PackedGraph (one per query).  Vertices, edges and half-edges of problem p stay
    contiguous (``prob_voff``, ``prob_eoff``); the library then keeps residuals, rho and the stop decision per
    problem and needs no communication between them."""
    voff = np.concatenate([[0], np.cumsum([g.nV for g in graphs])]).astype(np.int64)
    eoff = np.concatenate([[0], np.cumsum([g.nE for g in graphs])]).astype(np.int64)
    roff = np.concatenate([[0], np.cumsum([int(g.poly_off[-1]) for g in graphs])]).astype(np.int64)
    off = np.concatenate([[0]] + [g.poly_off[1:].astype(np.int64) + roff[i] for i, g in enumerate(graphs)])
    A = np.concatenate([g.polyA for g in graphs])
    b = np.concatenate([g.polyb for g in graphs])
    tail = np.concatenate([g.edge_tail.astype(np.int64) + voff[i] for i, g in enumerate(graphs)])
    head = np.concatenate([g.edge_head.astype(np.int64) + voff[i] for i, g in enumerate(graphs)])
    srcs = np.array([g.src + voff[i] if g.src >= 0 else -1 for i, g in enumerate(graphs)], dtype=np.int64)
    dsts = np.array([g.dst + voff[i] if g.dst >= 0 else -1 for i, g in enumerate(graphs)], dtype=np.int64)
    big = PackedGraph(off, A, b, tail, head, srcs, dsts)
    big.prob_voff = voff.astype(np.int32)
    big.prob_eoff = eoff.astype(np.int32)
    return big
