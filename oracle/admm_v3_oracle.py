"""ORACLE (test infrastructure, not product code) — numpy/fp64 restatement of the
reference's full-vertex-split ADMM, ``/root/reference/admm_solver_v3.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this file; the product path never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` replays this oracle
against the reference's own stored runs ``benchmark_data/admm_solver_v3_benchmark{1..4}.pkl``
(iteration count, every entry of ``pri_res_seq`` / ``dual_res_seq``, cost, ``y_v``)
exported to ``tests/golden/`` by ``tools/export_golden.py``.

What follows the reference line by line
---------------------------------------
* variable blocks and the consensus rows  ``admm_solver_v3.py:68-137``, ``:142-198``:
  for an edge e=(u,w) the rows are, for dim in range(n) ONLY (first point),
  z_u^e = z_u^{e,u}, z_u^e = z_u^{e,w}, z_w^e = z_w^{e,w}, z_w^e = z_w^{e,u}, then
  y_e^e = y_e^u, y_e^e = y_e^w.  Each endpoint v therefore owns 2n+1 = 5 consensus
  scalars per incident edge, stored here in *edge-canonical* order
  ``xc[h] = (copy of z_u[:n], copy of z_w[:n], copy of y_e)``.
* per-vertex program  ``:352-466`` (cost ``:380-413``, C1-C4 ``:416-440``, C5 ``:443-447``,
  C6 ``:450-456``, C7 ``:460-464``) — built literally (all 9+9d variables) and solved by
  the dense interior-point method in ``gcs_admm_b200.conic`` (the reference hands it to
  MOSEK through Drake, ``:490``; neither is available offline).
* edge averaging ``:543-562``; dual update ``:590-594``; residuals ``:597-602``;
  eps ``:605-614``; loop, rho adaptation and stop rule ``:621-733``; cost ``GCS_utils.py:184-211``.

Presolve (the reference relies on MOSEK's): constraints whose slack is zero on the
whole feasible set are removed before the interior-point solve, because they leave no
strict interior.  They are exactly: the incoming edges of 's' and outgoing edges of 't'
(flow forced to 0), all edges of a vertex with no live in- or out-edge, and for 's'/'t'
rows C2, C4, y<=1 (implied by C3 summed over the other live edges once y_v = 1).
The feasible set and the objective are unchanged.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import gcs_admm_b200  # noqa: E402,F401
from gcs_admm_b200.conic import solve_conic_qp  # noqa: E402

EDGE_PENALTY = 1e-4      # admm_solver_v3.py:388


class Params:
    """Literals of the reference's main loop (admm_solver_v3.py:621-651)."""
    rho0 = 1.0
    tau_incr = 2.0
    tau_decr = 2.0
    nu = 10.0
    frac = 0.1
    eps_abs = 1e-4
    eps_rel = 1e-3
    max_it = 1000


class VertexProgram:
    """Static part (G, E, cones) of one vertex's sub-problem; only the linear
    cost and the quadratic weight change between ADMM iterations."""

    def __init__(self, g, v):
        n = 2
        A = g.polyA[g.poly_off[v]:g.poly_off[v + 1]]
        b = g.polyb[g.poly_off[v]:g.poly_off[v + 1]]
        m = A.shape[0]
        hs = np.arange(g.he_off[v], g.he_off[v + 1])
        d = hs.shape[0]
        self.v, self.hs, self.d = v, hs, d
        is_s, is_t = (v == g.src), (v == g.dst)
        out = g.he_out[hs].astype(bool)
        # forced-zero half-edges
        zero = np.zeros(d, dtype=bool)
        if is_s:
            zero |= ~out
        if is_t:
            zero |= out
        live_in = int(np.sum(~out & ~zero))
        live_out = int(np.sum(out & ~zero))
        if (not is_s and live_in == 0) or (not is_t and live_out == 0):
            zero[:] = True
        self.zero = zero
        self.dead = bool(np.all(zero)) and not (is_s or is_t)
        if (is_s and live_out == 0) or (is_t and live_in == 0):
            raise ValueError("source has no outgoing edge / target has no incoming edge: infeasible")
        # variable layout: x(4) z(4) yv t | per half-edge: c0(4) c1(4) y
        X, Z, YV, T = 0, 4, 8, 9
        base = 10
        nvar = base + 9 * d
        self.nvar = nvar

        def C0(j): return base + 9 * j
        def C1(j): return base + 9 * j + 4
        def Y(j): return base + 9 * j + 8
        def OWN(j): return C0(j) if out[j] else C1(j)     # own copy: indexed by v itself
        self.idx = dict(X=X, Z=Z, YV=YV, T=T, C0=C0, C1=C1, Y=Y, OWN=OWN)
        terminal = is_s or is_t
        rows, rhs = [], []

        def add(coefs, r):
            row = np.zeros(nvar)
            for k, c in coefs:
                row[k] += c
            rows.append(row)
            rhs.append(r)

        # bounds on y_v (admm_solver_v3.py:366)
        if not terminal:
            add([(YV, -1.0)], 0.0)
            add([(YV, 1.0)], 1.0)
        for i in range(2):
            for r in range(m):
                # C1: A z_i <= y_v b                      (:420-422)
                add([(Z + 2 * i, A[r, 0]), (Z + 2 * i + 1, A[r, 1]), (YV, -b[r])], 0.0)
                if not terminal:
                    # C2: A (x_i - z_i) <= (1 - y_v) b    (:424-426)
                    add([(X + 2 * i, A[r, 0]), (X + 2 * i + 1, A[r, 1]),
                         (Z + 2 * i, -A[r, 0]), (Z + 2 * i + 1, -A[r, 1]), (YV, b[r])], b[r])
        for j in range(d):
            if zero[j]:
                continue
            add([(Y(j), -1.0)], 0.0)                         # y_e^v >= 0   (:377)
            if not terminal:
                add([(Y(j), 1.0)], 1.0)                      # y_e^v <= 1
            o = OWN(j)
            for i in range(2):
                for r in range(m):
                    # C3: A z^e_{v,i} <= y_e^v b           (:434-436)
                    add([(o + 2 * i, A[r, 0]), (o + 2 * i + 1, A[r, 1]), (Y(j), -b[r])], 0.0)
                    if not terminal:
                        # C4: A (x_i - z^e_{v,i}) <= (1 - y_e^v) b   (:438-440)
                        add([(X + 2 * i, A[r, 0]), (X + 2 * i + 1, A[r, 1]),
                             (o + 2 * i, -A[r, 0]), (o + 2 * i + 1, -A[r, 1]), (Y(j), b[r])], b[r])
        self.l = len(rows)
        # SOC (t ; z_1 - z_2):  s = h - G u  in Q^3     (:380-384)
        add([(T, -1.0)], 0.0)
        add([(Z, -1.0), (Z + 2, 1.0)], 0.0)
        add([(Z + 1, -1.0), (Z + 3, 1.0)], 0.0)
        self.G = np.array(rows)
        self.h = np.array(rhs)
        # equalities
        rows, rhs = [], []
        for j in range(d):
            for k in range(n):                                # C5 (:443-447)
                add([(C0(j) + n + k, 1.0), (C1(j) + k, -1.0)], 0.0)
            if zero[j]:
                o = OWN(j)
                for k in range(4):
                    add([(o + k, 1.0)], 0.0)
                add([(Y(j), 1.0)], 0.0)
        ds, dt = float(is_s), float(is_t)
        add([(YV, 1.0)] + [(Y(j), -1.0) for j in range(d) if not out[j]], ds)     # C6 (:454)
        add([(YV, 1.0)] + [(Y(j), -1.0) for j in range(d) if out[j]], dt)         # C6 (:456)
        for k in range(4):                                                        # C7 (:460-464)
            add([(Z + k, 1.0), (X + k, -ds)] + [(OWN(j) + k, -1.0) for j in range(d) if not out[j]], 0.0)
            add([(Z + k, 1.0), (X + k, -dt)] + [(OWN(j) + k, -1.0) for j in range(d) if out[j]], 0.0)
        self.E = np.array(rows)
        self.f = np.array(rhs)
        # quadratic selector: xc[j] = (c0[0:2], c1[0:2], y)
        self.sel = np.array([[C0(j), C0(j) + 1, C1(j), C1(j) + 1, Y(j)] for j in range(d)], dtype=int).reshape(-1)
        self.lin = np.zeros(nvar)
        self.lin[T] = 1.0
        for j in range(d):
            self.lin[Y(j)] = EDGE_PENALTY                                        # (:387-388)
        self.centroid = None

    def solve(self, rho, target, tol=1e-10):
        """argmin of the vertex program for consensus targets ``target`` (d,5)."""
        P = np.zeros((self.nvar, self.nvar))
        q = self.lin.copy()
        if self.d:
            P[self.sel, self.sel] = rho
            q[self.sel] -= rho * target.reshape(-1)
        res = solve_conic_qp(P, q, self.G, self.h, self.l, (3,), self.E, self.f, tol=tol)
        return res


class OracleADMM:
    """State and iteration of admm_solver_v3.py:621-733 on a PackedGraph ``g``."""

    def __init__(self, g, params=None, inner_tol=1e-10):
        self.g = g
        self.p = params or Params()
        self.inner_tol = inner_tol
        H, nE, nV = g.H, g.nE, g.nV
        self.xc = np.zeros((H, 5))
        self.mu = np.zeros((H, 5))
        self.z = np.zeros((nE, 5))
        self.x_v = np.zeros((nV, 4))
        self.z_v = np.zeros((nV, 4))
        self.y_v = np.zeros(nV)
        self.rho = float(self.p.rho0)
        self.rho_seq = [self.rho]
        self.pri_seq = [0.0]       # admm_solver_v3.py:638-639 (all-zero start)
        self.dual_seq = [0.0]
        self.it = 0
        self.opt = False
        self.progs = [VertexProgram(g, v) for v in range(nV)]
        self.cent = g.interior_points()
        self.inner_iters = 0
        self.n_x = 9 * nV + 18 * nE       # len(x_global), admm_solver_v3.py:345
        self.n_mu = 10 * nE               # len(mu_global), :349

    # --- x-update (:352-540) ------------------------------------------------
    def vertex_update(self):
        g = self.g
        for v, prog in enumerate(self.progs):
            hs = prog.hs
            if prog.d == 0 or prog.dead:
                # no flow can pass: every own variable is 0; the free "other copy"
                # first point of an incoming edge sits on its target.
                self.z_v[v] = 0.0
                self.y_v[v] = 0.0
                self.x_v[v] = np.r_[self.cent[v], self.cent[v]]
                for j, h in enumerate(hs):
                    e = g.he_edge[h]
                    tgt = self.z[e] + self.mu[h]
                    self.xc[h] = 0.0
                    if not g.he_out[h]:
                        self.xc[h, 0:2] = tgt[0:2]
                continue
            target = self.z[g.he_edge[hs]] + self.mu[hs]
            res = prog.solve(self.rho, target, self.inner_tol)
            self.inner_iters += res.iterations
            if res.status != "optimal" and max(res.pres, res.dres) > 1e-6:
                raise RuntimeError(f"oracle inner solve failed at vertex {v}: {res.status} "
                                   f"pres={res.pres:.2e} dres={res.dres:.2e} gap={res.gap:.2e}")
            u = res.u
            self.x_v[v] = u[0:4]
            self.z_v[v] = u[4:8]
            self.y_v[v] = u[8]
            self.xc[hs] = u[prog.sel].reshape(-1, 5)

    # --- z-update (:543-587) --------------------------------------------------
    def edge_update(self):
        g = self.g
        self.z_prev = self.z.copy()
        self.z = 0.5 * (self.xc[g.edge_he_tail] + self.xc[g.edge_he_head])

    def step(self):
        """One pass of the while-loop body, admm_solver_v3.py:655-733. Returns True on stop."""
        g, p = self.g, self.p
        self.it += 1
        it = self.it
        self.vertex_update()
        self.edge_update()
        r = self.z[g.he_edge] - self.xc                     # A x + B z - c, one row per (h, k)
        self.mu = self.mu + r                               # :594
        pri = float(np.sqrt(np.sum(r * r)))                 # :598
        dz = self.z - self.z_prev
        dual = self.rho * float(np.sqrt(2.0 * np.sum(dz * dz)))   # :602  ||A'B dz|| = sqrt2 ||dz||
        self.pri_seq.append(pri)
        self.dual_seq.append(dual)
        if pri >= p.nu * dual and it < p.frac * p.max_it:   # :703-708
            self.rho *= p.tau_incr
            self.mu /= p.tau_incr
        elif dual >= p.nu * pri and it < p.frac * p.max_it:
            self.rho *= 1.0 / p.tau_decr
            self.mu *= p.tau_incr
        self.rho_seq.append(self.rho)
        nAx = float(np.sqrt(np.sum(self.xc * self.xc)))     # :606-610
        nBz = float(np.sqrt(2.0 * np.sum(self.z * self.z)))
        eps_pri = np.sqrt(self.n_x) * p.eps_abs + p.eps_rel * max(nAx, nBz, 0.0)
        eps_dual = np.sqrt(self.n_mu) * p.eps_abs + p.eps_rel * float(np.sqrt(np.sum(self.mu * self.mu)))
        if pri < eps_pri and dual < eps_dual:               # :712
            self.opt = True
        return self.opt

    def run(self, max_it=None, stop=True):
        max_it = max_it or self.p.max_it
        while self.it < max_it:
            if self.step() and stop:
                break
        return self

    def cost(self):
        """GCS_utils.py:184-211 on the last iterates (admm_solver_v3.py:745-750)."""
        zv = self.z_v
        return float(np.sum(np.linalg.norm(zv[:, 0:2] - zv[:, 2:4], axis=1)) + EDGE_PENALTY * np.sum(self.z[:, 4]))
