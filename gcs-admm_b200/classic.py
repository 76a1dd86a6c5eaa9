"""Drake-free restatement of the reference's classic solver (``classic_solver.py``): the convex relaxation of the
GCS shortest-path MICP as ONE conic program, solved on the host by the dense interior-point method of
``conic.py``.  This is the CPU comparator the north star names for configurations where the reference's v3 is too
slow; it is not on the CUDA path.

Variables, in the reference's order of creation (``classic_solver.py:56-84``): per vertex ``x_v(2n) z_v(2n) y_v``,
per edge ``y_e``, per (vertex, incident edge) ``z^e_v(2n)``; plus one epigraph variable per vertex for the
``AddL2NormCost`` terms (``:90-96``).  Constraints C1-C7 (``:107-161``); ``0 <= y <= 1`` (``:66, :75``).

Presolve (the reference leaves it to MOSEK; an interior-point method needs a strictly feasible interior): edges
into 's' and out of 't' carry no flow (C6 with y <= 1), a non-terminal vertex without a live in- or out-edge
carries none either (cascading); their variables are fixed at 0.  For 's' / 't' C6-C7 give y_v = 1, z_v = x_v, so
C2 is the equality x_v = z_v again and C4, y_v <= 1 are implied by C3 summed over the other live edges: dropped.
``y_e <= 1`` and ``y_v >= 0`` are implied by C6 and are dropped everywhere.  Feasible set and optimum unchanged.
"""
from __future__ import annotations

import time

import numpy as np

from .conic import solve_conic_qp
from .graph import build_graph, delta
from .rounding import rounding

__all__ = ["solve_classic"]


def solve_classic(As, bs, n, *, graph=None, round_solution=True, seed=None, tol=1e-9, verbose=False, kkt=True):
    """Returns dict(cost, x_v_sol, z_v_sol, y_v_sol, y_e_sol, z_v_e_sol, solve_time, status[, final_cost,
    x_v_rounded, y_v_rounded, path]); ``cost`` is the relaxation optimum the reference prints as
    "Optimal Cost Pre-rounding" (``:208-209``)."""
    V0, E0, I_in0, I_out0 = graph if graph is not None else build_graph(As, bs)
    # ---- presolve: live edges / vertices
    live = {e for e in E0 if e[1] != 's' and e[0] != 't'}
    dead_v = set()
    changed = True
    while changed:
        changed = False
        for v in V0:
            if v in dead_v:
                continue
            has_in = any(e in live for e in I_in0[v]) or v == 's'
            has_out = any(e in live for e in I_out0[v]) or v == 't'
            if not (has_in and has_out):
                dead_v.add(v)
                for e in I_in0[v] + I_out0[v]:
                    if e in live:
                        live.discard(e); changed = True
    if 's' in dead_v or 't' in dead_v:
        raise ValueError("no path from 's' to 't' in the overlap graph")
    V = [v for v in V0 if v not in dead_v]
    E = [e for e in E0 if e in live]
    I_in = {v: [e for e in I_in0[v] if e in live] for v in V}
    I_out = {v: [e for e in I_out0[v] if e in live] for v in V}
    nV, nE, d2 = len(V), len(E), 2 * n
    vi = {v: i for i, v in enumerate(V)}
    ei = {e: i for i, e in enumerate(E)}
    ox, oz, oy = 0, d2 * nV, 2 * d2 * nV
    oye = oy + nV
    ozve = oye + nE
    pairs = [(v, e) for v in V for e in I_in[v] + I_out[v]]
    pi = {p: i for i, p in enumerate(pairs)}
    ot = ozve + d2 * len(pairs)
    nvar = ot + nV
    X = lambda v, c: ox + d2 * vi[v] + c          # noqa: E731
    Z = lambda v, c: oz + d2 * vi[v] + c          # noqa: E731
    Y = lambda v: oy + vi[v]                      # noqa: E731
    YE = lambda e: oye + ei[e]                    # noqa: E731
    ZE = lambda v, e, c: ozve + d2 * pi[(v, e)] + c   # noqa: E731

    rows, rhs = [], []

    def ineq(coefs, b):
        rows.append(coefs); rhs.append(b)

    for v in V:                                   # y_v <= 1 (non-terminals), y_e >= 0
        ineq({Y(v): 1.0}, 1.0 if v not in ('s', 't') else 2.0)      # terminals: y_v = 1 by C6; the slack row only keeps y_v in G
    for e in E:
        ineq({YE(e): -1.0}, 0.0)
    for v in V:
        A, b = np.asarray(As[v], float), np.asarray(bs[v], float).reshape(-1)
        term = v in ('s', 't')
        for i in range(2):
            for j in range(A.shape[0]):
                if term:                                                            # C1 + C2 with y_v = 1:  A x_i <= b  (z_v = x_v by C7)
                    ineq({X(v, i * n + k): A[j, k] for k in range(n)}, b[j])
                    continue
                c = {Z(v, i * n + k): A[j, k] for k in range(n)}                    # C1  A z_i <= y_v b
                c[Y(v)] = -b[j]
                ineq(c, 0.0)
                c = {X(v, i * n + k): A[j, k] for k in range(n)}                    # C2  A (x_i - z_i) <= (1 - y_v) b
                c.update({Z(v, i * n + k): -A[j, k] for k in range(n)})
                c[Y(v)] = b[j]
                ineq(c, b[j])
        for e in I_in[v] + I_out[v]:
            for i in range(2):
                for j in range(A.shape[0]):
                    c = {ZE(v, e, i * n + k): A[j, k] for k in range(n)}            # C3
                    c[YE(e)] = -b[j]
                    ineq(c, 0.0)
                    if term:
                        continue
                    c = {X(v, i * n + k): A[j, k] for k in range(n)}                # C4
                    c.update({ZE(v, e, i * n + k): -A[j, k] for k in range(n)})
                    c[YE(e)] = b[j]
                    ineq(c, b[j])
    l = len(rows)
    socs = []
    for v in V:                                   # (t_v ; z_v1 - z_v2) in SOC  <-  AddL2NormCost :90-96
        ineq({ot + vi[v]: -1.0}, 0.0)
        for k in range(n):
            ineq({Z(v, k): -1.0, Z(v, n + k): 1.0}, 0.0)
        socs.append(n + 1)
    import scipy.sparse as sp

    def assemble(dict_rows):
        ri = np.fromiter((r for r, c in enumerate(dict_rows) for _ in c), dtype=np.int64)
        ci = np.fromiter((k for c in dict_rows for k in c), dtype=np.int64)
        va = np.fromiter((val for c in dict_rows for val in c.values()), dtype=np.float64)
        return sp.csr_matrix((va, (ri, ci)), shape=(len(dict_rows), nvar))      # duplicate (row, column) entries are summed
    G = assemble(rows)
    h = np.array(rhs)

    erows, f = [], []
    for e in E:                                   # C5
        v, w = e
        for k in range(n):
            erows.append({ZE(v, e, n + k): 1.0, ZE(w, e, k): -1.0}); f.append(0.0)
    c6_out_row = {}
    for v in V:
        ds, dt = delta('s', v), delta('t', v)
        c = {Y(v): 1.0}; c.update({YE(e): -1.0 for e in I_in[v]}); erows.append(c); f.append(float(ds))      # C6
        c6_out_row[v] = len(erows)
        c = {Y(v): 1.0}; c.update({YE(e): -1.0 for e in I_out[v]}); erows.append(c); f.append(float(dt))
        for k in range(d2):                       # C7
            c = {Z(v, k): 1.0}; c.update({ZE(v, e, k): -1.0 for e in I_in[v]})
            if ds:
                c[X(v, k)] = c.get(X(v, k), 0.0) - 1.0
            erows.append(c); f.append(0.0)
            c = {Z(v, k): 1.0}; c.update({ZE(v, e, k): -1.0 for e in I_out[v]})
            if dt:
                c[X(v, k)] = c.get(X(v, k), 0.0) - 1.0
            erows.append(c); f.append(0.0)
    # C6 summed over the vertices of one connected component is 0 = 0 (every edge is one in- and one out-edge of it): one
    # redundant row per component — drop the "out" flow row of the component's first vertex.  (Up to a few hundred vertices
    # this is cross-checked against a pivoted QR of the dense matrix.)
    from scipy.sparse.csgraph import connected_components
    adj = sp.csr_matrix((np.ones(nE), ([vi[e[0]] for e in E], [vi[e[1]] for e in E])), shape=(nV, nV)) if nE else sp.csr_matrix((nV, nV))
    ncomp, comp = connected_components(adj, directed=False)
    first = {}
    for v in V:
        first.setdefault(comp[vi[v]], v)
    drop = {c6_out_row[v] for v in first.values()}
    keep = [k for k in range(len(erows)) if k not in drop]
    Em = assemble(erows)[keep]
    f = [f[k] for k in keep]
    if nV <= 300:
        from scipy.linalg import qr
        _, R, _ = qr(Em.toarray().T, mode="economic", pivoting=True)
        dg = np.abs(np.diag(R))
        if int(np.sum(dg > 1e-10 * dg[0])) != Em.shape[0]:
            raise RuntimeError("equality rows of the classic program are not independent after the analytic reduction")
    q = np.zeros(nvar)
    q[oye:oye + nE] = 1e-4                        # :99-100
    q[ot:] = 1.0
    t0 = time.time()
    res = solve_conic_qp(None, q, G, h, l, socs, Em, np.array(f), tol=tol, max_iter=100, augmented=kkt)
    solve_time = time.time() - t0
    u = res.u
    out = dict(status=res.status, iterations=res.iterations, solve_time=solve_time, V=V0, E=E0,
               residuals=dict(gap=float(res.gap), pres=float(res.pres), dres=float(res.dres)))
    zero = np.zeros(d2)
    out["x_v_sol"] = {v: (u[X(v, 0):X(v, 0) + d2].copy() if v in vi else zero.copy()) for v in V0}
    out["z_v_sol"] = {v: (u[Z(v, 0):Z(v, 0) + d2].copy() if v in vi else zero.copy()) for v in V0}

    def readable(val):                            # :181-185, :191-201
        if abs(val) < 1e-6:
            return 0
        if abs(val) > 1 - 1e-6:
            return 1
        return float(val)

    out["y_v_sol"] = {v: (readable(u[Y(v)]) if v in vi else 0) for v in V0}
    out["y_e_sol"] = {e: (readable(u[YE(e)]) if e in ei else 0) for e in E0}
    out["z_v_e_sol"] = {(v, e): (u[ZE(v, e, 0):ZE(v, e, 0) + d2].copy() if (v, e) in pi else zero.copy())
                        for v in V0 for e in I_in0[v] + I_out0[v]}
    out["cost"] = float(q @ u)
    if round_solution:
        fc, xr, yr, path = rounding(out["y_e_sol"], V0, E0, I_out0, As, bs, n, rng=seed, return_path=True)
        out.update(final_cost=fc, x_v_rounded=xr, y_v_rounded=yr, path=path)
    return out
