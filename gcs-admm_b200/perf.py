"""Host-side tables of the `perf` mode of K1 (``csrc/vertex_perf.cuh``).

* ``class_inverse``: for a vertex class (type, #live in-edges, #live out-edges) the matrix
  ``K1 = N'(S'S + kappa M'M)N`` of the v-step and its inverse.  ``N`` is the null-space map of the equalities
  C6/C7 (reference ``admm_solver_v3.py:450-464``), ``S`` selects the consensus scalars that carry the rho-quadratic
  (``:390-413``), ``M`` maps the variables to the (point, flow) pairs of C1-C4 (``:416-440``) and to ``z_1 - z_2``.
  All three have 0/+-1 entries and depend on the class only — not on the polygon.
* ``cone_table``: the rays / face normals of every region's perspective cone for the exact 3-D projection.

The Python functions ``_forward`` / ``_pair_values`` restate ``gcs_forward`` / ``gcs_pair_values`` of the kernels.
"""
from __future__ import annotations

import numpy as np

from .graph import _vertices_batch

NCORE, UX, UT, UZ, UYV = 10, 0, 4, 5, 9


def _uw(j):
    return NCORE + 5 * j


def _forward(v, d, jstar, prim, term, affine):
    u = np.zeros(NCORE + 5 * d)
    u[UX:UX + 4] = v[0:4]
    u[UT] = v[4]
    zy = np.zeros(5)
    other = np.zeros(5)
    if term:
        zy[0:4] = v[0:4]
        zy[4] = 1.0 if affine else 0.0
    for jj in range(d - 1):
        j = jj if jj < jstar else jj + 1
        w = v[5 + 5 * jj: 10 + 5 * jj]
        u[_uw(j):_uw(j) + 5] = w
        if prim[j]:
            other += w
        else:
            zy += w
    u[UZ:UZ + 5] = zy
    u[_uw(jstar):_uw(jstar) + 5] = zy - other
    return u


def _pair_values(u, d, term):
    npair = 4 * (d + 1)
    pv = np.zeros(3 * npair + 2)
    for blk in range(d + 1):
        for i in range(2):
            po = (_uw(blk) if blk < d else UZ) + 2 * i
            yo = _uw(blk) + 4 if blk < d else UYV
            xo = UX + 2 * i
            for fam in range(1 if term else 2):
                o = 3 * ((blk * 2 + i) * 2 + fam)
                if fam:
                    pv[o:o + 3] = (u[xo] - u[po], u[xo + 1] - u[po + 1], 1.0 - u[yo])
                else:
                    pv[o:o + 3] = (u[po], u[po + 1], u[yo])
    pv[3 * npair] = u[UZ] - u[UZ + 2]
    pv[3 * npair + 1] = u[UZ + 1] - u[UZ + 3]
    return pv


def class_pattern(vtype, din, dout):
    """(d, out flags, prim flags, jstar, term) of a class, in the kernels' half-edge order (in-edges then out-edges)."""
    d = din + dout
    out = [0] * din + [1] * dout
    term = vtype != 0
    prim = [1] * d if vtype == 2 else list(out)
    jstar = max(j for j in range(d) if prim[j])
    return d, out, prim, jstar, term


def class_inverse(vtype, din, dout, kappa):
    d, out, prim, jstar, term = class_pattern(vtype, din, dout)
    n, nu = 5 * d, NCORE + 5 * d
    N = np.zeros((nu, n))
    for k in range(n):
        e = np.zeros(n); e[k] = 1.0
        N[:, k] = _forward(e, d, jstar, prim, term, False)
    pv0 = _pair_values(np.zeros(nu), d, term)
    M = np.zeros((pv0.shape[0], nu))
    for k in range(nu):
        e = np.zeros(nu); e[k] = 1.0
        M[:, k] = _pair_values(e, d, term) - pv0
    S = np.zeros(nu)
    for j in range(d):
        o = _uw(j)
        S[o:o + 2] = 1.0
        if out[j]:
            S[o + 2:o + 4] = 1.0
        S[o + 4] = 1.0
    K1 = N.T @ (np.diag(S) + kappa * M.T @ M) @ N
    K1[4, :] = 0.0; K1[:, 4] = 0.0; K1[4, 4] = 1.0          # the epigraph variable t is unused in this mode
    w = np.linalg.eigvalsh(K1)
    if w[0] <= 1e-12 * w[-1]:
        raise ValueError(f"singular v-step matrix for class {(vtype, din, dout)}")
    Kinv = np.linalg.inv(K1)
    Kinv[4, :] = 0.0; Kinv[:, 4] = 0.0
    return np.ascontiguousarray(Kinv)


def cone_table(g):
    """Per region: polygon vertices (counter-clockwise) with the unit outward normal of the cone face spanned by the
    rays through vertex k and k+1, 1/|r_k|^2 (r_k = (V_k, 1)) and the two in-plane sector normals of that face
    (12 doubles per polygon vertex).  Vectorised over regions with the same vertex count."""
    verts, cnt = g.polygon_vertices_batch() if hasattr(g, "polygon_vertices_batch") else _vertices_batch(g.poly_off.astype(np.int64), g.polyA, g.polyb)
    nV = g.nV
    # duplicates (redundant rows meeting in one vertex) are rare: handle those regions one by one
    per = [None] * nV
    for k in np.unique(cnt):
        idx = np.nonzero(cnt == k)[0]
        P = verts[idx, :k]                                            # (g, k, 2)
        c = P.mean(axis=1, keepdims=True)
        ang = np.arctan2(P[..., 1] - c[..., 1], P[..., 0] - c[..., 0])
        order = np.argsort(ang, axis=1)
        P = np.take_along_axis(P, order[..., None], axis=1)
        nxt = np.roll(P, -1, axis=1)
        dup = np.max(np.abs(P - nxt), axis=2) <= 1e-12 * np.maximum(1.0, np.abs(P).max(axis=2))
        clean = ~np.any(dup, axis=1)
        R = np.concatenate([P, np.ones(P.shape[:2] + (1,))], axis=2)
        Nn = np.cross(R, np.roll(R, -1, axis=1))
        nrm = np.linalg.norm(Nn, axis=2, keepdims=True)
        Nn = Nn / np.where(nrm > 0, nrm, 1.0)
        inner = np.concatenate([c, np.ones((c.shape[0], 1, 1))], axis=2)
        flip = np.sum(Nn * inner, axis=2) > 0
        Nn = np.where(flip[..., None], -Nn, Nn)
        Rn = np.roll(R, -1, axis=1)
        # in-plane sector tests of face k: p = a r + b s with a, b >= 0  <=>  p.ma >= 0 and p.mb >= 0, where
        # ma = n x s (zero on s, positive on r) and mb = r x n (zero on r, positive on s), n = outward unit normal
        ma = np.cross(Nn, Rn); mb = np.cross(R, Nn)
        ma *= np.sign(np.sum(ma * R, axis=2, keepdims=True)); mb *= np.sign(np.sum(mb * Rn, axis=2, keepdims=True))
        rec = np.concatenate([P, Nn, 1.0 / np.sum(R * R, axis=2, keepdims=True), ma, mb], axis=2)      # (g, k, 12)
        for gi, v in enumerate(idx):
            per[v] = rec[gi] if clean[gi] else None
    for v in range(nV):
        if per[v] is None:
            per[v] = _cone_one(verts[v, :cnt[v]])
    off = np.zeros(nV + 1, dtype=np.int64)
    np.cumsum([p.shape[0] for p in per], out=off[1:])
    return off.astype(np.int32), np.ascontiguousarray(np.vstack(per))


def _cone_one(P):
    keep = []
    for p in P:
        if not any(np.max(np.abs(p - q)) <= 1e-12 * max(1.0, np.abs(p).max()) for q in keep):
            keep.append(p)
    P = np.array(keep)
    c = P.mean(axis=0)
    P = P[np.argsort(np.arctan2(P[:, 1] - c[1], P[:, 0] - c[0]))]
    R = np.hstack([P, np.ones((P.shape[0], 1))])
    Nn = np.cross(R, np.roll(R, -1, axis=0))
    Nn /= np.linalg.norm(Nn, axis=1, keepdims=True)
    flip = (Nn @ np.array([c[0], c[1], 1.0])) > 0
    Nn[flip] *= -1.0
    Rn = np.roll(R, -1, axis=0)
    ma = np.cross(Nn, Rn); mb = np.cross(R, Nn)
    ma *= np.sign(np.sum(ma * R, axis=1, keepdims=True)); mb *= np.sign(np.sum(mb * Rn, axis=1, keepdims=True))
    return np.hstack([P, Nn, 1.0 / np.sum(R * R, axis=1, keepdims=True), ma, mb])


def perf_tables(g, kappa=1.0):
    owner = g.he_owner.astype(np.int64)
    live = (g.he_flags & 2) == 0
    outm = (g.he_flags & 1) == 1
    din = np.bincount(owner[live & ~outm], minlength=g.nV)
    dout = np.bincount(owner[live & outm], minlength=g.nV)
    keys = {}
    vclass = np.full(g.nV, -1, dtype=np.int32)
    mats, koff, pos = [], [], 0
    code = g.vtype.astype(np.int64) * 1000000 + din * 1000 + dout
    for cd in np.unique(code[g.vtype != 3]):
        key = (int(cd // 1000000), int((cd // 1000) % 1000), int(cd % 1000))
        keys[key] = len(keys)
        Kinv = class_inverse(*key, kappa)
        mats.append(Kinv.reshape(-1)); koff.append(pos); pos += Kinv.size
        vclass[(code == cd) & (g.vtype != 3)] = keys[key]
    cone_off, cone = cone_table(g)
    return dict(vclass=vclass, class_koff=np.array(koff, dtype=np.int32), kinv=np.concatenate(mats), cone_off=cone_off,
                cone=cone, classes=keys, kappa=float(kappa))


def local_tables(T, lp):
    """Tables of one rank's share of the graph (``partition.LocalProblem``): the class inverses are global, the
    per-vertex class ids and cone records are sliced in the rank's vertex order."""
    lv = np.asarray(lp.global_vertices, dtype=np.int64)
    off = T["cone_off"].astype(np.int64)
    cnt = off[lv + 1] - off[lv]
    loff = np.zeros(lv.shape[0] + 1, dtype=np.int64)
    np.cumsum(cnt, out=loff[1:])
    rec = T["cone"].reshape(-1, 12)
    idx = np.repeat(off[lv] - loff[:-1], cnt) + np.arange(int(loff[-1]))
    out = dict(T)
    out.update(vclass=T["vclass"][lv].copy(), cone_off=loff.astype(np.int32), cone=np.ascontiguousarray(rec[idx]))
    return out
