"""Small dense conic-QP interior-point solver (numpy, fp64) — HOST side only.

    minimize    1/2 u'Pu + q'u
    subject to  G u + s = h,  s in K = R_+^l x Q^{q_1} x ... x Q^{q_k}
                E u = f

Used by the host for the path-restricted SOCPs of the rounding step (reference
``GCS_utils.py:17-89`` solves those with Drake ``Solve``); it is not on the
per-iteration ADMM path, which is the CUDA library's job.

Method: primal-dual path following with Nesterov-Todd scaling and Mehrotra's
predictor-corrector (the standard scheme for symmetric cones).  The Newton
system is reduced to  (P + G' W^-1 W^-T G) du + E' dy = rhs ;  E du = rhs  and
solved by Cholesky + Schur complement on the equality block.
"""
from __future__ import annotations

import numpy as np

__all__ = ["solve_conic_qp", "ConicResult"]


class ConicResult:
    __slots__ = ("u", "s", "z", "y", "iterations", "gap", "pres", "dres", "status", "obj")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class _Cone:
    """Product cone bookkeeping: first ``l`` coordinates are R_+, then SOCs."""

    def __init__(self, l, socs):
        self.l = int(l)
        self.socs = [int(q) for q in socs]
        self.dim = self.l + sum(self.socs)
        self.degree = self.l + len(self.socs)
        self.blocks = []
        o = self.l
        for q in self.socs:
            self.blocks.append((o, o + q))
            o += q

    def identity(self):
        e = np.zeros(self.dim)
        e[:self.l] = 1.0
        for a, _ in self.blocks:
            e[a] = 1.0
        return e

    def prod(self, u, v):
        w = np.empty(self.dim)
        w[:self.l] = u[:self.l] * v[:self.l]
        for a, b in self.blocks:
            w[a] = u[a:b] @ v[a:b]
            w[a + 1:b] = u[a] * v[a + 1:b] + v[a] * u[a + 1:b]
        return w

    def div(self, lam, v):
        """x with lam o x = v."""
        x = np.empty(self.dim)
        x[:self.l] = v[:self.l] / lam[:self.l]
        for a, b in self.blocks:
            l0, l1 = lam[a], lam[a + 1:b]
            det = l0 * l0 - l1 @ l1
            x0 = (l0 * v[a] - l1 @ v[a + 1:b]) / det
            x[a] = x0
            x[a + 1:b] = (v[a + 1:b] - x0 * l1) / l0
        return x

    def max_step(self, lam, d):
        """sup{alpha : lam + alpha d in K} as 1/t with t = max(0, -min eig(lam^-1/2-scaled d));
        returns t (0 means unbounded)."""
        t = 0.0
        if self.l:
            t = max(t, float(np.max(-d[:self.l] / lam[:self.l])))
        for a, b in self.blocks:
            l0, l1 = lam[a], lam[a + 1:b]
            nrm = np.sqrt(max(l0 * l0 - l1 @ l1, 1e-300))
            lb0, lb1 = l0 / nrm, l1 / nrm
            # rho = (1/nrm) * H(lb)^-1-type scaling of d  (CVXOPT misc.max_step)
            c0 = (lb0 * d[a] - lb1 @ d[a + 1:b])
            c1 = d[a + 1:b] - (c0 + d[a]) / (lb0 + 1.0) * lb1
            c0 /= nrm
            c1 = c1 / nrm
            t = max(t, float(np.linalg.norm(c1) - c0))
        return t


class _Scaling:
    """Nesterov-Todd scaling W with  W z = W^-T s = lam."""

    def __init__(self, cone, s, z):
        self.cone = cone
        l = cone.l
        self.d = np.sqrt(s[:l] / z[:l])          # W for the LP part
        self.beta, self.w = [], []
        lam = np.empty(cone.dim)
        lam[:l] = np.sqrt(s[:l] * z[:l])
        for a, b in cone.blocks:
            sa, za = s[a:b], z[a:b]
            sn = np.sqrt(max(sa[0] ** 2 - sa[1:] @ sa[1:], 1e-300))
            zn = np.sqrt(max(za[0] ** 2 - za[1:] @ za[1:], 1e-300))
            sb, zb = sa / sn, za / zn
            gamma = np.sqrt((1.0 + sb @ zb) / 2.0)
            w = np.empty(b - a)
            w[0] = (sb[0] + zb[0]) / (2 * gamma)
            w[1:] = (sb[1:] - zb[1:]) / (2 * gamma)
            self.beta.append(np.sqrt(sn / zn))
            self.w.append(w)
        self.lam = lam
        lam[:] = self.apply(z)                    # LP part recomputed identically

    def _soc(self, k, x, inverse):
        w, beta = self.w[k], self.beta[k]
        w0, w1 = w[0], w[1:]
        if inverse:
            w1 = -w1
        y = np.empty_like(x)
        t = w1 @ x[1:]
        y[0] = w0 * x[0] + t
        y[1:] = x[1:] + (x[0] + t / (1.0 + w0)) * w1
        return y / beta if inverse else y * beta

    def apply(self, x):
        """W x"""
        y = np.empty(self.cone.dim)
        l = self.cone.l
        y[:l] = self.d * x[:l]
        for k, (a, b) in enumerate(self.cone.blocks):
            y[a:b] = self._soc(k, x[a:b], False)
        return y

    def apply_inv(self, x):
        """W^-1 x  (W symmetric, so also W^-T x)"""
        y = np.empty(self.cone.dim)
        l = self.cone.l
        y[:l] = x[:l] / self.d
        for k, (a, b) in enumerate(self.cone.blocks):
            y[a:b] = self._soc(k, x[a:b], True)
        return y

    def gram_inv(self, G):
        """G' (W'W)^-1 G  and the operator  v -> (W'W)^-1 v  applied to G's rows."""
        l = self.cone.l
        WiG = np.empty_like(G)
        WiG[:l] = G[:l] / self.d[:, None]
        for k, (a, b) in enumerate(self.cone.blocks):
            for j in range(G.shape[1]):
                WiG[a:b, j] = self._soc(k, G[a:b, j], True)
        return WiG


def solve_conic_qp(P, q, G, h, l, socs=(), E=None, f=None, *, tol=1e-10, max_iter=80, reg=1e-13, centrality=1e-2,
                   augmented=False, static_reg=1e-9):
    """Solve the conic QP above.  ``G`` rows: first ``l`` linear inequalities, then
    the SOC blocks of sizes ``socs`` (each block (t; w) means t >= ||w||).

    ``augmented=True``: the Newton systems are solved on the sparse quasi-definite augmented matrix
    ``[P + dI, E', G'; E, -dI, 0; G, 0, -W'W]`` (sparse LU, static regularisation ``d`` removed by iterative
    refinement) instead of the dense normal equations — for large, sparse, degenerate programs (``classic.py``)."""
    n = q.shape[0]
    cone = _Cone(l, socs)
    if augmented:
        import scipy.sparse as sp
        from scipy.sparse.linalg import splu
        Ps = sp.csc_matrix((n, n)) if P is None else sp.csc_matrix(P)
        Gs = sp.csc_matrix(G)
        Es = sp.csc_matrix((0, n)) if E is None else sp.csc_matrix(E)
        if E is None:
            f = np.zeros(0)
        G, E, P = Gs, Es, Ps            # matrix-vector products below work unchanged on sparse matrices
        mz = cone.dim
    else:
        P = np.zeros((n, n)) if P is None else np.asarray(P, float)
        if E is None:
            E = np.zeros((0, n))
            f = np.zeros(0)
    p = E.shape[0]
    e = cone.identity()

    def kkt_factor_aug(W):
        if W is None:
            W2 = sp.identity(mz, format="csc")
        else:
            dlin = W.d * W.d
            blocks = [sp.diags(dlin)] if cone.l else []
            for k, (a, b) in enumerate(cone.blocks):
                q_ = b - a
                B = np.empty((q_, q_))
                for j in range(q_):
                    ej = np.zeros(q_); ej[j] = 1.0
                    B[:, j] = W._soc(k, W._soc(k, ej, False), False)
                blocks.append(sp.csc_matrix(B))
            W2 = sp.block_diag(blocks, format="csc")
        K = sp.bmat([[Ps + static_reg * sp.identity(n), Es.T if p else None, Gs.T],
                     [Es if p else None, -static_reg * sp.identity(p) if p else None, None],
                     [Gs, None, -W2]], format="csc") if p else \
            sp.bmat([[Ps + static_reg * sp.identity(n), Gs.T], [Gs, -W2]], format="csc")
        if not np.all(np.isfinite(K.data)):
            raise np.linalg.LinAlgError('non-finite KKT matrix')
        try:
            return splu(K)      # COLAMD + partial pivoting: measured more robust here than the symmetric-mode orderings
        except RuntimeError as ex:
            raise np.linalg.LinAlgError(str(ex))

    def kkt_solve_once_aug(fac, W, bx, by, bz):
        if augmented == "reduced":
            lu, Wi2 = fac
            r = bx + Gs.T @ (Wi2 @ bz)
            sol = lu.solve(np.concatenate([r, by]))
            du = sol[:n]
            return du, sol[n:], Wi2 @ (Gs @ du - bz)
        sol = fac.solve(np.concatenate([bx, by, bz]))
        return sol[:n], sol[n:n + p], sol[n + p:]

    def kkt_factor_red(W):
        """``augmented="reduced"``: the slack block is eliminated, leaving the quasi-definite
        ``[P + G'W^-2 G + dI, E'; E, -dI]`` of dimension n + p — an order of magnitude smaller than the 3 x 3 form when the
        program has many more inequality rows than variables (``classic.py``: ~6 rows per variable)."""
        if W is None:
            Wi2 = sp.identity(mz, format="csr")
        else:
            blocks = [sp.diags(1.0 / (W.d * W.d))] if cone.l else []
            for k, (a, b) in enumerate(cone.blocks):
                q_ = b - a
                B = np.empty((q_, q_))
                for j in range(q_):
                    ej = np.zeros(q_); ej[j] = 1.0
                    B[:, j] = W._soc(k, W._soc(k, ej, True), True)
                blocks.append(sp.csr_matrix(B))
            Wi2 = sp.block_diag(blocks, format="csr")
        H = Ps + (Gs.T @ Wi2 @ Gs) + static_reg * sp.identity(n)
        K = sp.bmat([[H, Es.T], [Es, -static_reg * sp.identity(p)]], format="csc") if p else sp.csc_matrix(H)
        if not np.all(np.isfinite(K.data)):
            raise np.linalg.LinAlgError('non-finite KKT matrix')
        try:
            return splu(K), Wi2
        except RuntimeError as ex:
            raise np.linalg.LinAlgError(str(ex))

    def kkt_factor(W):
        if augmented == "reduced":
            return kkt_factor_red(W)
        if augmented:
            return kkt_factor_aug(W)
        WiG = W.gram_inv(G) if W is not None else G
        H0 = P + WiG.T @ WiG
        # static regularisation relative to the matrix scale, raised until the
        # factorisation succeeds; its effect is removed by iterative refinement
        delta = reg
        if not np.all(np.isfinite(H0)):
            raise np.linalg.LinAlgError('non-finite KKT matrix')
        while True:
            try:
                L = np.linalg.cholesky(H0 + delta * np.eye(n))
                break
            except np.linalg.LinAlgError:
                delta *= 100.0
                if delta > 1e-3 * max(1.0, float(np.max(np.diag(H0)))):
                    raise
        if p:
            Li_Et = np.linalg.solve(L, E.T)
            S = Li_Et.T @ Li_Et
            ds = reg
            while True:             # same lifting as for H0; iterative refinement removes its effect
                try:
                    LS = np.linalg.cholesky(S + ds * np.eye(p))
                    break
                except np.linalg.LinAlgError:
                    ds *= 100.0
                    if ds > 1e-4 * max(1.0, float(np.max(np.diag(S)))):
                        raise
        else:
            Li_Et = LS = None
        return L, Li_Et, LS, H0

    def reduced_solve(fac, r, by):
        L, Li_Et, LS, _ = fac
        Lr = np.linalg.solve(L, r)
        if p:
            rhs = Li_Et.T @ Lr - by
            dy = np.linalg.solve(LS.T, np.linalg.solve(LS, rhs))
            Lr = Lr - Li_Et @ dy
        else:
            dy = np.zeros(0)
        return np.linalg.solve(L.T, Lr), dy

    def kkt_solve_once(fac, W, bx, by, bz):
        if augmented:
            return kkt_solve_once_aug(fac, W, bx, by, bz)
        if W is not None:
            t = W.apply_inv(W.apply_inv(bz))
        else:
            t = bz
        r = bx + G.T @ t
        du, dy = reduced_solve(fac, r, by)
        v = G @ du - bz
        dz = W.apply_inv(W.apply_inv(v)) if W is not None else v
        return du, dy, dz

    def kkt_solve(fac, W, bx, by, bz, refine=4):
        """[P E' G'; E 0 0; G 0 -W'W] (du,dy,dz) = (bx,by,bz), with iterative
        refinement on the full (unreduced, unregularised) system."""
        du, dy, dz = kkt_solve_once(fac, W, bx, by, bz)
        nb = np.linalg.norm(bx) + np.linalg.norm(by) + np.linalg.norm(bz) + 1e-300
        for _ in range(refine):
            e1 = bx - (P @ du + G.T @ dz + (E.T @ dy if p else 0.0))
            e2 = by - E @ du if p else np.zeros(0)
            WWdz = W.apply(W.apply(dz)) if W is not None else dz
            e3 = bz - (G @ du - WWdz)
            if np.linalg.norm(e1) + np.linalg.norm(e2) + np.linalg.norm(e3) <= 1e-15 * nb:
                break
            cu, cy, cz = kkt_solve_once(fac, W, e1, e2, e3)
            du, dy, dz = du + cu, dy + cy, dz + cz
        return du, dy, dz

    # initial point (W = I)
    fac = kkt_factor(None)
    u, y, z = kkt_solve(fac, None, -q, f, h)
    s = -z.copy()
    ts = cone.max_step(e, s)      # s + t e in K  <=>  t >= ts  (lam = e)
    if ts >= -1e-8 * max(1.0, np.linalg.norm(s)):
        s = s + (1.0 + ts) * e
    tz = cone.max_step(e, z)
    if tz >= -1e-8 * max(1.0, np.linalg.norm(z)):
        z = z + (1.0 + tz) * e

    resx0 = max(1.0, np.linalg.norm(q))
    resz0 = max(1.0, np.linalg.norm(h))
    resy0 = max(1.0, np.linalg.norm(f)) if p else 1.0
    status = "max_iter"
    it = 0
    gap = pres = dres = np.inf
    for it in range(max_iter + 1):
        rx = P @ u + q + G.T @ z + (E.T @ y if p else 0.0)
        ry = E @ u - f if p else np.zeros(0)
        rz = G @ u + s - h
        gap = float(s @ z)
        pcost = 0.5 * u @ P @ u + q @ u
        dcost = pcost + (y @ ry if p else 0.0) + z @ rz - gap
        pres = max(np.linalg.norm(ry) / resy0, np.linalg.norm(rz) / resz0)
        dres = np.linalg.norm(rx) / resx0
        if pcost < 0:
            relgap = gap / -pcost
        elif dcost > 0:
            relgap = gap / dcost
        else:
            relgap = np.inf
        if pres <= tol and dres <= tol and (gap <= tol or relgap <= tol):
            status = "optimal"
            break
        if it == max_iter:
            break
        W = _Scaling(cone, s, z)
        lam = W.lam
        try:
            fac = kkt_factor(W)
        except np.linalg.LinAlgError:
            status = "numerical"
            break
        mu = gap / cone.degree
        lamsq = cone.prod(lam, lam)

        def direction(ds, sigma_scale):
            # ds is the rhs of  lam o (W dz + W^-T ds) = ds
            t = cone.div(lam, ds)
            bx = -sigma_scale * rx
            by = -sigma_scale * ry
            bz = -sigma_scale * rz - W.apply(t)
            du, dy, dz = kkt_solve(fac, W, bx, by, bz)
            # scaled ds:  W^-T ds = t - W dz
            Wdz = W.apply(dz)
            Wds = t - Wdz
            return du, dy, dz, Wds, Wdz

        du, dy, dz, Wds, Wdz = direction(-lamsq, 1.0)
        ts = cone.max_step(lam, Wds)
        tz = cone.max_step(lam, Wdz)
        t = max(ts, tz)
        alpha = 1.0 if t <= 0 else min(1.0, 1.0 / t)
        sigma = (1.0 - alpha) ** 3
        ds = -lamsq - cone.prod(Wds, Wdz) + sigma * mu * e
        du, dy, dz, Wds, Wdz = direction(ds, 1.0 - sigma)
        ts = cone.max_step(lam, Wds)
        tz = cone.max_step(lam, Wdz)
        t = max(ts, tz)
        alpha = 1.0 if t <= 0 else min(1.0, 0.99 / t)
        dsv = W.apply(Wds)                  # ds = W^T (W^-T ds), W symmetric
        # stay in a wide neighbourhood of the central path: no complementarity
        # product may fall below gamma * (average product) after the step
        for _ in range(20):
            sn, zn = s + alpha * dsv, z + alpha * dz
            prods = [sn[:l] * zn[:l]] + [np.array([sn[a:b] @ zn[a:b]]) for a, b in cone.blocks]
            prods = np.concatenate(prods)
            if np.min(prods) >= centrality * np.sum(prods) / cone.degree:
                break
            alpha *= 0.7
        u = u + alpha * du
        y = y + alpha * dy
        z = zn
        s = sn
    obj = 0.5 * u @ P @ u + q @ u
    return ConicResult(u=u, s=s, z=z, y=y, iterations=it, gap=gap, pres=pres, dres=dres,
                       status=status, obj=float(obj))
