"""PROTOTYPE (numpy, CPU) for a round-2 `perf` mode of K1: replace the exact interior-point solve of a vertex
program by K warm-started iterations of an operator-splitting scheme whose steps are all closed-form:

  * every (point, flow) pair of the program lies in the perspective cone K_P = {(p, h): A p <= h b} of the
    vertex's polygon (C1: (z_i, y_v), C2: (x_i - z_i, 1 - y_v), C3: (a_i, y_e), C4: (x_i - a_i, 1 - y_e));
    the pairs are 0/+-1 linear images of the variables, so the splitting  c = M u + m0,  c in prod K_P  has a
    linear system that depends only on (degree pattern, rho, sigma) — not on the polygon;
  * projection onto K_P in R^3 is exact in O(m): interior / a face / a ray through a polygon vertex / the apex;
  * the path-length term |z_1 - z_2| is a block soft-threshold.

The script measures (1) inner iterations needed to match the exact solution, (2) whether the outer ADMM of
admm_solver_v3 still converges to the classic optimum when its x-update is only K such iterations.
Not product code; nothing imports it.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import utils  # noqa: E402,F401
from admm_v3_oracle import OracleADMM, VertexProgram, EDGE_PENALTY  # noqa: E402
from conftest import load_golden  # noqa: E402
from gcs_admm_b200.graph import pack_graph, polygon_vertices  # noqa: E402


class ConeProjector:
    """Exact projection onto K_P = {(p, h) in R^3 : A p <= h b} for a bounded polygon P."""

    def __init__(self, A, b):
        self.A, self.b = A, b
        self.N = np.hstack([A, -b[:, None]])                 # rows n_k: n_k . (p, h) <= 0
        self.Nn = self.N / np.linalg.norm(self.N, axis=1, keepdims=True)
        V = polygon_vertices(A, b)
        self.R = np.hstack([V, np.ones((V.shape[0], 1))])    # rays through the polygon's vertices
        self.R2 = np.sum(self.R * self.R, axis=1)

    def __call__(self, c):
        if np.all(self.N @ c <= 1e-15):
            return c
        best, bd = np.zeros(3), float(c @ c)
        for k in range(self.Nn.shape[0]):                    # faces
            q = c - (self.Nn[k] @ c) * self.Nn[k]
            if np.all(self.N @ q <= 1e-12 * max(1.0, np.abs(q).max())):
                d = float((c - q) @ (c - q))
                if d < bd:
                    best, bd = q, d
        tau = np.maximum(0.0, self.R @ c / self.R2)         # rays
        Q = tau[:, None] * self.R
        d = np.sum((Q - c) ** 2, axis=1)
        k = int(np.argmin(d))
        if d[k] < bd:
            best, bd = Q[k], float(d[k])
        return best


class SplitVertex:
    """Operator-splitting form of one literal vertex program (layout of oracle VertexProgram)."""

    def __init__(self, g, v, prog, sigma=1.0):
        self.prog, self.kappa, self.sigma = prog, sigma, sigma
        A = g.polyA[g.poly_off[v]:g.poly_off[v + 1]]
        b = g.polyb[g.poly_off[v]:g.poly_off[v + 1]]
        self.proj = ConeProjector(A, b)
        ix = prog.idx
        X, Z, YV = ix["X"], ix["Z"], ix["YV"]
        n = prog.nvar
        terminal = (v == g.src) or (v == g.dst)
        rows, off = [], []

        def pair(pt_plus, pt_minus, h_idx, h_sign, h_const):
            for c in range(2):
                r = np.zeros(n); r[pt_plus + c] = 1.0
                if pt_minus is not None:
                    r[pt_minus + c] = -1.0
                rows.append(r); off.append(0.0)
            r = np.zeros(n); r[h_idx] = h_sign
            rows.append(r); off.append(h_const)
        for i in range(2):
            pair(Z + 2 * i, None, YV, 1.0, 0.0)
            if not terminal:
                pair(X + 2 * i, Z + 2 * i, YV, -1.0, 1.0)
        for j in range(prog.d):
            if prog.zero[j]:
                continue
            o, y = ix["OWN"](j), ix["Y"](j)
            for i in range(2):
                pair(o + 2 * i, None, y, 1.0, 0.0)
                if not terminal:
                    pair(X + 2 * i, o + 2 * i, y, -1.0, 1.0)
        self.npairs = len(rows) // 3
        for c in range(2):                                   # path-length term  w = z_1 - z_2
            r = np.zeros(n); r[Z + c] = 1.0; r[Z + 2 + c] = -1.0
            rows.append(r); off.append(0.0)
        self.M, self.m0 = np.array(rows), np.array(off)
        # the epigraph variable t of the literal program is unused here: pin it
        E = np.vstack([prog.E, np.eye(n)[ix["T"]][None, :]])
        f = np.concatenate([prog.f, [0.0]])
        self.E, self.f = E, f
        self.c = np.zeros(self.M.shape[0]); self.lam = np.zeros(self.M.shape[0])
        self.u = np.zeros(n)
        self._kkt_rho = None

    def _factor(self, rho):
        if self._kkt_rho is not None:
            self.lam *= self._kkt_rho / rho          # scaled dual of the splitting follows sigma = kappa * rho
        self.sigma = self.kappa * rho
        n, p = self.prog.nvar, self.E.shape[0]
        P = np.zeros((n, n)); P[self.prog.sel, self.prog.sel] = rho
        K = np.zeros((n + p, n + p))
        K[:n, :n] = P + self.sigma * self.M.T @ self.M + 1e-12 * np.eye(n)
        K[:n, n:] = self.E.T; K[n:, :n] = self.E
        K[n:, n:] = -1e-12 * np.eye(p)
        self.Kinv = np.linalg.inv(K)
        self._kkt_rho = rho

    def iterate(self, rho, target, K, alpha=1.6):
        if self._kkt_rho != rho:
            self._factor(rho)
        n = self.prog.nvar
        q = np.zeros(n)
        for j in range(self.prog.d):
            q[self.prog.idx["Y"](j)] = EDGE_PENALTY
        q[self.prog.sel] -= rho * target.reshape(-1)
        for _ in range(K):
            rhs = np.concatenate([-q + self.sigma * self.M.T @ (self.c - self.lam - self.m0), self.f])
            self.u = (self.Kinv @ rhs)[:n]
            Mu = self.M @ self.u + self.m0
            Mr = alpha * Mu + (1 - alpha) * self.c
            v = Mr + self.lam
            cn = np.empty_like(v)
            for k in range(self.npairs):
                cn[3 * k:3 * k + 3] = self.proj(v[3 * k:3 * k + 3])
            w = v[-2:]; nw = np.linalg.norm(w)
            cn[-2:] = max(0.0, 1.0 - 1.0 / (self.sigma * nw)) * w if nw > 0 else 0.0
            self.lam = self.lam + Mr - cn
            self.c = cn
        return self.u


def run(name, K, iters, sigma=1.0, compare_every=0):
    sigma = float(os.environ.get('KAPPA', sigma))
    if name.startswith('grid'):
        from gcs_admm_b200.generator import grid_packed_graph
        g = grid_packed_graph(int(name[4:])); d = {'classic_cost': np.nan, 'v3_cost': np.nan}
    else:
        As, bs, n, d, keys = load_golden(name)
        g = pack_graph(As, bs)
    o = OracleADMM(g)                   # used for its graph bookkeeping, edge / dual / residual arithmetic
    splits = {}
    for v, prog in enumerate(o.progs):
        if not (prog.d == 0 or prog.dead):
            splits[v] = SplitVertex(g, v, prog, sigma)

    def inexact_vertex_update():
        for v, prog in enumerate(o.progs):
            hs = prog.hs
            if v not in splits:
                o.z_v[v] = 0.0; o.y_v[v] = 0.0
                for h in hs:
                    tgt = o.z[g.he_edge[h]] + o.mu[h]
                    o.xc[h] = 0.0
                    if not g.he_out[h]:
                        o.xc[h, 0:2] = tgt[0:2]
                continue
            target = o.z[g.he_edge[hs]] + o.mu[hs]
            u = splits[v].iterate(o.rho, target, K)
            o.x_v[v] = u[0:4]; o.z_v[v] = u[4:8]; o.y_v[v] = u[8]
            o.xc[hs] = u[prog.sel].reshape(-1, 5)
    o.vertex_update = inexact_vertex_update
    t0 = time.time()
    for it in range(1, iters + 1):
        o.step()
        if it % max(1, iters // 10) == 0:
            print(f"  K={K} it {it:5d} pri {o.pri_seq[-1]:.3e} dual {o.dual_seq[-1]:.3e} cost {o.cost():.6f} "
                  f"(classic {float(d['classic_cost']):.6f}, v3@stop {float(d['v3_cost']):.6f})  [{time.time() - t0:.0f}s]")
    return o


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "benchmark1"
    Ks = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [5, 20]
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    for K in Ks:
        print(f"{name}: outer ADMM with K = {K} inner splitting iterations per x-update")
        run(name, K, iters)
