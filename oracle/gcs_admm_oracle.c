/*
 * ORACLE — test infrastructure, NOT product code.
 *
 * Plain-C, fp64, CPU restatement of the reference's full-vertex-split ADMM
 * (/root/reference/admm_solver_v3.py).  Only tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline / --impl reference legs may load this library; the
 * product path (gcs-admm_b200/csrc) never does.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py replays it against the
 * reference's stored runs benchmark_data/admm_solver_v3_benchmark{1..4}.pkl
 * (exported to tests/golden/ by tools/export_golden.py) and cross-checks the
 * per-vertex solves against the literal numpy restatement oracle/admm_v3_oracle.py.
 *
 * What follows the reference (file:line = /root/reference/admm_solver_v3.py):
 *   consensus rows / variable blocks        :68-137, :142-198
 *   per-vertex program (x-update)           :352-466   (solved by MOSEK at :490)
 *   edge averaging (z-update)               :543-562
 *   dual update                             :590-594
 *   residuals, eps_pri, eps_dual            :597-614
 *   loop order, rho adaptation, stop rule   :621-733
 *   cost                                    GCS_utils.py:184-211
 *
 * The per-vertex program is solved in its REDUCED form (the reductions are exact,
 * see DESIGN.md "vertex program"): for an incident edge e=(u,w) of vertex v the
 * copy indexed by the OTHER endpoint enters only through its first point (the
 * consensus quadratic) and C5; so with a = own copy (a1,a2 in R^2), y = y_e^v:
 *     outgoing (v=u): other.first == a2 (C5)  -> quadratic on a1, a2, y
 *     incoming (v=w): other.first is free     -> equals its target; quadratic on a1, y
 * Unknowns: x_v(4), z_v(4), y_v, t (epigraph of ||z_v1 - z_v2||), and (a1,a2,y) per
 * live half-edge.  Interior-point method: primal-feasible start, Mehrotra
 * predictor-corrector, Nesterov-Todd scaling for the one second-order cone, dense
 * Cholesky of the reduced KKT system + Schur complement on the equalities.
 *
 * Presolve (the reference leaves this to MOSEK): half-edges whose flow is forced to 0
 * (incoming edges of 's', outgoing edges of 't', every edge of a vertex lacking a live
 * in- or out-edge) are fixed at 0; for 's'/'t' (y_v = 1, z_v = x_v) rows C2, C4 are
 * dropped because C3 summed over the other live edges implies them.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EDGE_PENALTY 1e-4 /* admm_solver_v3.py:388 */
#define NC_MAX 10
#define NE_MAX 10

typedef struct {
    int nV, nE;
    int *poly_off; double *polyA, *polyb;
    int *he_off, *he_edge, *he_out, *edge_he_tail, *edge_he_head;
    int src, dst;
    double *cent;
    /* params (admm_solver_v3.py:621-651) */
    double rho0, tau_incr, tau_decr, nu, frac, eps_abs, eps_rel;
    int max_it;
    double inner_tol; int inner_max_iter;
    /* state */
    double *xc, *mu, *z, *zprev, *x_v, *z_v, *y_v;
    double rho; int it, opt, diverged;
    double *rho_seq, *pri_seq, *dual_seq; int hist_cap;
    long inner_iters; int inner_fail;
    unsigned char *he_zero; /* forced-zero half-edges */
    unsigned char *vtype;   /* 0 generic, 1 source, 2 target, 3 dead */
} Oracle;

/* ---------------------------------------------------------------- dense helpers */
/* In-place lower Cholesky (row-major) with pivot lifting: a pivot that falls below the
 * rounding noise of its own cancellation (a direction the data cannot resolve in fp64)
 * is lifted to that noise level, which bounds the Newton step along it instead of
 * amplifying garbage.  Returns the number of lifted pivots. */
static int g_verbose = 0;
void gcso_set_verbose(int v) { g_verbose = v; }
static int chol(double *M, int n) {
    int lifted = 0;
    for (int j = 0; j < n; ++j) {
        double d0 = M[j * n + j], sq = 0.0;
        for (int k = 0; k < j; ++k) sq += M[j * n + k] * M[j * n + k];
        double d = d0 - sq, noise = 64.0 * 2.2e-16 * (fabs(d0) + sq) + 1e-300;
        if (!(d > noise)) { if (g_verbose) printf("      lift pivot %d of %d: d0=%.3e d=%.3e noise=%.3e\n", j, n, d0, d, noise); d = noise; lifted++; }
        d = sqrt(d);
        M[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = M[i * n + j];
            for (int k = 0; k < j; ++k) s -= M[i * n + k] * M[j * n + k];
            M[i * n + j] = s / d;
        }
    }
    return lifted;
}
static void chol_fwd(const double *L, int n, double *x) {
    for (int i = 0; i < n; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * x[k];
        x[i] = s / L[i * n + i];
    }
}
static void chol_bwd(const double *L, int n, double *x) {
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
}

/* ---------------------------------------------------------------- vertex program */
typedef struct { int idx[5]; double c[5]; int nnz; double h; } Row;

typedef struct {
    int nvar, nc, ne, nrow, d, term;
    int X, Z, YV, T;     /* core offsets (YV = -1 for terminals) */
    Row *rows;           /* linear inequality rows: s = h - g.u >= 0 */
    double *E, *f;       /* ne x nvar, ne */
    double *Pd, *q;      /* diagonal quadratic, linear cost */
    double *u0;          /* strictly feasible start */
} VProg;

static void row_set(Row *r, double h) { r->nnz = 0; r->h = h; }
static void row_add(Row *r, int i, double c) { r->idx[r->nnz] = i; r->c[r->nnz] = c; r->nnz++; }

/* Builds the reduced program of one vertex.  `out[j]` = 1 if live half-edge j is outgoing.
 * tgt[j*5..] = consensus targets in edge-canonical order (z_u[:2], z_w[:2], y). */
static VProg *vprog_build(int m, const double *A, const double *b, int d, const int *out, int term,
                          const double *cent, double rho, const double *tgt) {
    VProg *p = (VProg *)calloc(1, sizeof(VProg));
    p->d = d; p->term = term;
    if (term) { p->X = 0; p->Z = 0; p->YV = -1; p->T = 4; p->nc = 5; p->ne = 5; }
    else      { p->X = 0; p->Z = 4; p->YV = 8;  p->T = 9; p->nc = 10; p->ne = 10; }
    int nc = p->nc;
    p->nvar = nc + 5 * d;
    int nvar = p->nvar;
    int cap = 1 + 4 * m + d * (1 + 4 * m);
    p->rows = (Row *)calloc(cap, sizeof(Row));
    int nr = 0;
    Row *r;
    if (!term) { r = &p->rows[nr++]; row_set(r, 1.0); row_add(r, p->YV, 1.0); } /* y_v <= 1 (:366) */
    for (int i = 0; i < 2; ++i)
        for (int k = 0; k < m; ++k) {
            r = &p->rows[nr++];
            if (term) { /* C1 with y_v = 1, z_v = x_v (:420-422) */
                row_set(r, b[k]); row_add(r, p->X + 2 * i, A[2 * k]); row_add(r, p->X + 2 * i + 1, A[2 * k + 1]);
            } else {
                row_set(r, 0.0); /* C1: A z_i <= y_v b */
                row_add(r, p->Z + 2 * i, A[2 * k]); row_add(r, p->Z + 2 * i + 1, A[2 * k + 1]); row_add(r, p->YV, -b[k]);
                r = &p->rows[nr++];
                row_set(r, b[k]); /* C2: A (x_i - z_i) <= (1 - y_v) b (:424-426) */
                row_add(r, p->X + 2 * i, A[2 * k]); row_add(r, p->X + 2 * i + 1, A[2 * k + 1]);
                row_add(r, p->Z + 2 * i, -A[2 * k]); row_add(r, p->Z + 2 * i + 1, -A[2 * k + 1]); row_add(r, p->YV, b[k]);
            }
        }
    for (int j = 0; j < d; ++j) {
        int base = nc + 5 * j, Y = base + 4;
        r = &p->rows[nr++]; row_set(r, 0.0); row_add(r, Y, -1.0); /* y_e^v >= 0 (:377) */
        for (int i = 0; i < 2; ++i)
            for (int k = 0; k < m; ++k) {
                r = &p->rows[nr++]; row_set(r, 0.0); /* C3: A a_i <= y b (:434-436) */
                row_add(r, base + 2 * i, A[2 * k]); row_add(r, base + 2 * i + 1, A[2 * k + 1]); row_add(r, Y, -b[k]);
                if (!term) {
                    r = &p->rows[nr++]; row_set(r, b[k]); /* C4: A (x_i - a_i) <= (1 - y) b (:438-440) */
                    row_add(r, p->X + 2 * i, A[2 * k]); row_add(r, p->X + 2 * i + 1, A[2 * k + 1]);
                    row_add(r, base + 2 * i, -A[2 * k]); row_add(r, base + 2 * i + 1, -A[2 * k + 1]); row_add(r, Y, b[k]);
                }
            }
    }
    p->nrow = nr;
    /* equalities: C6 (:450-456), C7 (:460-464) after eliminating the forced zeros */
    p->E = (double *)calloc((size_t)p->ne * nvar, sizeof(double));
    p->f = (double *)calloc(p->ne, sizeof(double));
    if (term) {
        for (int k = 0; k < 4; ++k) p->E[k * nvar + p->X + k] = -1.0;
        for (int j = 0; j < d; ++j) {
            int base = nc + 5 * j;
            for (int k = 0; k < 4; ++k) p->E[k * nvar + base + k] = 1.0;
            p->E[4 * nvar + base + 4] = 1.0;
        }
        p->f[4] = 1.0;
    } else {
        for (int k = 0; k < 4; ++k) { p->E[k * nvar + p->Z + k] = -1.0; p->E[(4 + k) * nvar + p->Z + k] = -1.0; }
        p->E[8 * nvar + p->YV] = -1.0; p->E[9 * nvar + p->YV] = -1.0;
        for (int j = 0; j < d; ++j) {
            int base = nc + 5 * j, g = out[j] ? 1 : 0;
            for (int k = 0; k < 4; ++k) p->E[(4 * g + k) * nvar + base + k] = 1.0;
            p->E[(8 + g) * nvar + base + 4] = 1.0;
        }
    }
    /* objective (:380-413) */
    p->Pd = (double *)calloc(nvar, sizeof(double));
    p->q = (double *)calloc(nvar, sizeof(double));
    p->q[p->T] = 1.0;
    for (int j = 0; j < d; ++j) {
        int base = nc + 5 * j;
        const double *t = tgt + 5 * j;
        const double *T1 = out[j] ? t : t + 2; /* own first point */
        p->Pd[base] = p->Pd[base + 1] = rho; p->q[base] = -rho * T1[0]; p->q[base + 1] = -rho * T1[1];
        if (out[j]) { p->Pd[base + 2] = p->Pd[base + 3] = rho; p->q[base + 2] = -rho * t[2]; p->q[base + 3] = -rho * t[3]; }
        p->Pd[base + 4] = rho; p->q[base + 4] = EDGE_PENALTY - rho * t[4];
    }
    /* strictly feasible start */
    p->u0 = (double *)calloc(nvar, sizeof(double));
    int din = 0, dout = 0;
    for (int j = 0; j < d; ++j) { if (out[j]) dout++; else din++; }
    double eta = term ? 1.0 : 0.5;
    for (int j = 0; j < d; ++j) {
        int base = nc + 5 * j;
        double y = eta / (double)(out[j] ? dout : din);
        p->u0[base] = p->u0[base + 2] = y * cent[0];
        p->u0[base + 1] = p->u0[base + 3] = y * cent[1];
        p->u0[base + 4] = y;
    }
    p->u0[p->X] = p->u0[p->X + 2] = cent[0]; p->u0[p->X + 1] = p->u0[p->X + 3] = cent[1];
    if (!term) {
        p->u0[p->Z] = p->u0[p->Z + 2] = eta * cent[0]; p->u0[p->Z + 1] = p->u0[p->Z + 3] = eta * cent[1];
        p->u0[p->YV] = eta;
    }
    p->u0[p->T] = 1.0;
    return p;
}
static void vprog_free(VProg *p) { free(p->rows); free(p->E); free(p->f); free(p->Pd); free(p->q); free(p->u0); free(p); }

/* second-order cone (dimension 3) helpers */
static double jnorm2(const double *s) { double n1 = hypot(s[1], s[2]); return (s[0] - n1) * (s[0] + n1); }
typedef struct { double w[3], beta, lam[3]; } NT;
static void nt_apply(const NT *S, const double *x, double *y, int inverse) {
    double w0 = S->w[0], w1 = inverse ? -S->w[1] : S->w[1], w2 = inverse ? -S->w[2] : S->w[2];
    double t = w1 * x[1] + w2 * x[2];
    double c = x[0] + t / (1.0 + w0);
    double y0 = w0 * x[0] + t, y1 = x[1] + c * w1, y2 = x[2] + c * w2;
    double sc = inverse ? 1.0 / S->beta : S->beta;
    y[0] = y0 * sc; y[1] = y1 * sc; y[2] = y2 * sc;
}
static void nt_build(NT *S, const double *s, const double *z) {
    double sn = sqrt(fmax(jnorm2(s), 1e-300)), zn = sqrt(fmax(jnorm2(z), 1e-300));
    double sb[3] = {s[0] / sn, s[1] / sn, s[2] / sn}, zb[3] = {z[0] / zn, z[1] / zn, z[2] / zn};
    double gamma = sqrt(0.5 * (1.0 + sb[0] * zb[0] + sb[1] * zb[1] + sb[2] * zb[2]));
    S->w[0] = (sb[0] + zb[0]) / (2 * gamma); S->w[1] = (sb[1] - zb[1]) / (2 * gamma); S->w[2] = (sb[2] - zb[2]) / (2 * gamma);
    S->beta = sqrt(sn / zn);
    nt_apply(S, z, S->lam, 0);
}
static void soc_prod(const double *a, const double *b, double *c) {
    c[0] = a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; c[1] = a[0] * b[1] + b[0] * a[1]; c[2] = a[0] * b[2] + b[0] * a[2];
}
static void soc_div(const double *lam, const double *v, double *x) { /* lam o x = v */
    double det = jnorm2(lam);
    double x0 = (lam[0] * v[0] - lam[1] * v[1] - lam[2] * v[2]) / det;
    x[0] = x0; x[1] = (v[1] - x0 * lam[1]) / lam[0]; x[2] = (v[2] - x0 * lam[2]) / lam[0];
}
static double soc_max_step(const double *lam, const double *dd) { /* t with lam + a d in K iff a <= 1/t */
    double nrm = sqrt(fmax(jnorm2(lam), 1e-300));
    double l0 = lam[0] / nrm, l1 = lam[1] / nrm, l2 = lam[2] / nrm;
    double c0 = l0 * dd[0] - l1 * dd[1] - l2 * dd[2];
    double f = (c0 + dd[0]) / (l0 + 1.0);
    double c1 = dd[1] - f * l1, c2 = dd[2] - f * l2;
    return (hypot(c1, c2) - c0) / nrm;
}

typedef struct { int iters; int status; double gap, dres, pres; } IpmInfo;

/* Null-space parametrisation of the equalities:  u = N v + up.
 * generic : z = sum_in a, y_v = sum_in y, and the LAST live out-edge block is
 *           w* = sum_in w - sum_{other out} w          v = [x(4), t, w_j (j != j*)]
 * terminal: x = sum a, y* = 1 - sum_{j != j*} y         v = [t, w_j (j != j*), a*(4)]
 * Every direction of v that the equalities tie to an out-edge inherits that edge's
 * rho-curvature, so the reduced Hessian stays well conditioned. */
typedef struct { int nred; double *N, *up; } NullSpace;

static NullSpace nullspace_build(const VProg *p, const int *out) {
    NullSpace ns;
    int nvar = p->nvar, nc = p->nc, d = p->d;
    int jstar = -1;
    for (int j = 0; j < d; ++j) if (p->term || out[j]) jstar = j;
    ns.nred = p->term ? 5 * d : 5 + 5 * (d - 1);
    ns.N = (double *)calloc((size_t)nvar * ns.nred, sizeof(double));
    ns.up = (double *)calloc(nvar, sizeof(double));
    int nred = ns.nred;
#define NN(i, k) ns.N[(size_t)(i) * nred + (k)]
    if (!p->term) {
        for (int k = 0; k < 4; ++k) NN(p->X + k, k) = 1.0;
        NN(p->T, 4) = 1.0;
        int col = 5;
        for (int j = 0; j < d; ++j) {
            if (j == jstar) continue;
            int base = nc + 5 * j, bs = nc + 5 * jstar;
            for (int k = 0; k < 5; ++k) {
                NN(base + k, col + k) = 1.0;
                NN(bs + k, col + k) = out[j] ? -1.0 : 1.0;          /* w* = sum_in - sum_other_out */
                if (!out[j]) { if (k < 4) NN(p->Z + k, col + k) = 1.0; else NN(p->YV, col + k) = 1.0; }
            }
            col += 5;
        }
    } else {
        NN(p->T, 0) = 1.0;
        int col = 1, bs = nc + 5 * jstar;
        for (int j = 0; j < d; ++j) {
            if (j == jstar) continue;
            int base = nc + 5 * j;
            for (int k = 0; k < 5; ++k) NN(base + k, col + k) = 1.0;
            for (int k = 0; k < 4; ++k) NN(p->X + k, col + k) = 1.0;
            NN(bs + 4, col + 4) = -1.0;
            col += 5;
        }
        for (int k = 0; k < 4; ++k) { NN(bs + k, col + k) = 1.0; NN(p->X + k, col + k) = 1.0; }
        ns.up[bs + 4] = 1.0;
    }
#undef NN
    return ns;
}

/* Solves the vertex program.  u (nvar) receives the solution.  status 0 = converged. */
static IpmInfo vprog_solve(const VProg *p, const int *out, double tol, int max_iter, double *u) {
    const int nvar = p->nvar, nr = p->nrow;
    const double gamma_nb = 1e-5;   /* width of the central-path neighbourhood */
    const double loqo_c = 0.02;     /* weight of the centrality-aware floor on sigma */
    IpmInfo info = {0, 1, 0, 0, 0};
    NullSpace ns = nullspace_build(p, out);
    const int n = ns.nred;
    /* reduced rows (dense), reduced objective, reduced SOC map */
    double *Gr = (double *)calloc((size_t)nr * n, sizeof(double)), *hr = (double *)malloc(sizeof(double) * nr);
    int *nzs = (int *)malloc(sizeof(int) * (size_t)nr * n), *nzc = (int *)calloc(nr, sizeof(int));
    for (int r = 0; r < nr; ++r) {
        const Row *R = &p->rows[r];
        double h = R->h;
        for (int k = 0; k < R->nnz; ++k) {
            h -= R->c[k] * ns.up[R->idx[k]];
            const double *Nrow = ns.N + (size_t)R->idx[k] * n;
            for (int c = 0; c < n; ++c) Gr[(size_t)r * n + c] += R->c[k] * Nrow[c];
        }
        hr[r] = h;
        for (int c = 0; c < n; ++c) if (Gr[(size_t)r * n + c] != 0.0) nzs[(size_t)r * n + nzc[r]++] = c;
    }
    double *Pr = (double *)calloc((size_t)n * n, sizeof(double)), *qr = (double *)calloc(n, sizeof(double));
    for (int i = 0; i < nvar; ++i) {
        const double *Ni = ns.N + (size_t)i * n;
        double g = p->q[i] + p->Pd[i] * ns.up[i];
        for (int a = 0; a < n; ++a) {
            if (Ni[a] == 0.0) continue;
            qr[a] += Ni[a] * g;
            if (p->Pd[i] != 0.0) for (int c = 0; c < n; ++c) Pr[(size_t)a * n + c] += p->Pd[i] * Ni[a] * Ni[c];
        }
    }
    double *Br = (double *)calloc((size_t)3 * n, sizeof(double));
    for (int c = 0; c < n; ++c) {
        Br[c] = ns.N[(size_t)p->T * n + c];
        Br[n + c] = ns.N[(size_t)p->Z * n + c] - ns.N[(size_t)(p->Z + 2) * n + c];
        Br[2 * n + c] = ns.N[(size_t)(p->Z + 1) * n + c] - ns.N[(size_t)(p->Z + 3) * n + c];
    }
    /* start: v0 with N v0 + up = u0 (u0 satisfies the equalities) */
    double *v = (double *)calloc(n, sizeof(double));
    for (int c = 0; c < n; ++c) { /* each column of N has a unit entry on an independent variable */
        for (int i = 0; i < nvar; ++i) {
            if (ns.N[(size_t)i * n + c] == 1.0) {
                int single = 1;
                for (int c2 = 0; c2 < n; ++c2) if (c2 != c && ns.N[(size_t)i * n + c2] != 0.0) { single = 0; break; }
                if (single) { v[c] = p->u0[i] - ns.up[i]; break; }
            }
        }
    }
    double *zl = (double *)malloc(sizeof(double) * nr), *sl = (double *)malloc(sizeof(double) * nr);
    double *dsl = (double *)malloc(sizeof(double) * nr), *dzl = (double *)malloc(sizeof(double) * nr);
    double *dsa = (double *)malloc(sizeof(double) * nr), *dza = (double *)malloc(sizeof(double) * nr);
    double *H = (double *)malloc(sizeof(double) * n * n), *L = (double *)malloc(sizeof(double) * n * n);
    double *rx = (double *)malloc(sizeof(double) * n), *rhs = (double *)malloc(sizeof(double) * n);
    double *dv = (double *)malloc(sizeof(double) * n), *e1 = (double *)malloc(sizeof(double) * n), *cv = (double *)malloc(sizeof(double) * n);
    double zq[3], sq[3];
#define SLACKS()                                                                              \
    do {                                                                                      \
        for (int r_ = 0; r_ < nr; ++r_) {                                                     \
            double s_ = hr[r_];                                                               \
            const double *g_ = Gr + (size_t)r_ * n; const int *z_ = nzs + (size_t)r_ * n;     \
            for (int k_ = 0; k_ < nzc[r_]; ++k_) s_ -= g_[z_[k_]] * v[z_[k_]];                \
            sl[r_] = s_;                                                                      \
        }                                                                                     \
        for (int a_ = 0; a_ < 3; ++a_) {                                                      \
            double s_ = 0.0;                                                                  \
            for (int c_ = 0; c_ < n; ++c_) s_ += Br[a_ * n + c_] * v[c_];                     \
            sq[a_] = s_;                                                                      \
        }                                                                                     \
    } while (0)
    SLACKS();
    double smean = 0.0;
    for (int r = 0; r < nr; ++r) smean += sl[r];
    smean /= nr;
    double mu0 = smean;
    for (int r = 0; r < nr; ++r) zl[r] = mu0 / sl[r];
    { double det = jnorm2(sq); zq[0] = mu0 * sq[0] / det; zq[1] = -mu0 * sq[1] / det; zq[2] = -mu0 * sq[2] / det; }
    double qn = 1.0;
    for (int i = 0; i < n; ++i) qn = fmax(qn, fabs(qr[i]));
    const int deg = nr + 1;
    int it, best_it = 0;
    double best_merit = 1e300, best_gap = 0, best_dres = 0;
    double *vbest = (double *)malloc(sizeof(double) * n);
    memcpy(vbest, v, sizeof(double) * n);
    for (it = 0; it <= max_iter; ++it) {
        SLACKS();
        double gap = sq[0] * zq[0] + sq[1] * zq[1] + sq[2] * zq[2];
        for (int r = 0; r < nr; ++r) gap += sl[r] * zl[r];
        /* rx = P v + q + G' z - B' zq */
        for (int a = 0; a < n; ++a) {
            double s = qr[a] - Br[a] * zq[0] - Br[n + a] * zq[1] - Br[2 * n + a] * zq[2];
            for (int c = 0; c < n; ++c) s += Pr[(size_t)a * n + c] * v[c];
            rx[a] = s;
        }
        for (int r = 0; r < nr; ++r) {
            const double *g = Gr + (size_t)r * n; const int *zi = nzs + (size_t)r * n;
            for (int k = 0; k < nzc[r]; ++k) rx[zi[k]] += g[zi[k]] * zl[r];
        }
        double dres = 0.0;
        for (int i = 0; i < n; ++i) dres = fmax(dres, fabs(rx[i]));
        dres /= qn;
        info.iters = it; info.gap = gap; info.dres = dres; info.pres = 0.0;
        if (g_verbose) printf("  ipm it %2d gap %.3e dres %.3e sq=(%.2e %.2e %.2e) zq=(%.2e %.2e %.2e)\n", it, gap, dres, sq[0], sq[1], sq[2], zq[0], zq[1], zq[2]);
        if (!(gap == gap) || !(dres == dres)) { info.status = 2; break; }
        if (dres <= 10.0 * tol && gap <= tol) { info.status = 0; break; }
        { double merit = fmax(gap / tol, dres / (10.0 * tol));
          if (merit < best_merit) { best_merit = merit; memcpy(vbest, v, sizeof(double) * n); best_gap = gap; best_dres = dres; }
          else if (gap <= tol && it >= best_it + 4) { info.status = 5; break; } /* stalled at the fp64 noise floor */
          if (merit <= best_merit) best_it = it; }
        if (it == max_iter) break;
        double mu = gap / deg;
        NT nt; nt_build(&nt, sq, zq);
        /* H = P + G' D G + B' W^-2 B */
        memcpy(H, Pr, sizeof(double) * n * n);
        for (int r = 0; r < nr; ++r) {
            const double *g = Gr + (size_t)r * n; const int *zi = nzs + (size_t)r * n;
            double D = zl[r] / sl[r];
            for (int a = 0; a < nzc[r]; ++a) {
                double ga = D * g[zi[a]];
                for (int c = 0; c < nzc[r]; ++c) H[(size_t)zi[a] * n + zi[c]] += ga * g[zi[c]];
            }
        }
        {
            double Wi2[3][3];
            for (int c = 0; c < 3; ++c) {
                double e[3] = {0, 0, 0}, y1[3], y2[3]; e[c] = 1.0;
                nt_apply(&nt, e, y1, 1); nt_apply(&nt, y1, y2, 1);
                for (int a = 0; a < 3; ++a) Wi2[a][c] = y2[a];
            }
            for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) {
                double w = Wi2[a][c];
                for (int i = 0; i < n; ++i) {
                    double bi = Br[a * n + i] * w;
                    if (bi == 0.0) continue;
                    for (int j = 0; j < n; ++j) H[(size_t)i * n + j] += bi * Br[c * n + j];
                }
            }
        }
        memcpy(L, H, sizeof(double) * n * n);
        for (int i = 0; i < n; ++i) L[(size_t)i * n + i] += 1e-14;
        int lifted = chol(L, n);
        double dsq_s[3], dzq_s[3], dsq_a[3], dzq_a[3], tq[3];
        double sigma = 0.0, alpha = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
            double dsoc[3];
            soc_prod(nt.lam, nt.lam, dsoc);
            dsoc[0] = -dsoc[0]; dsoc[1] = -dsoc[1]; dsoc[2] = -dsoc[2];
            if (pass == 1) {
                double cr[3]; soc_prod(dsq_a, dzq_a, cr);
                dsoc[0] += sigma * mu - cr[0]; dsoc[1] -= cr[1]; dsoc[2] -= cr[2];
            }
            soc_div(nt.lam, dsoc, tq);
            double scale = 1.0 - sigma;
            for (int i = 0; i < n; ++i) rhs[i] = -scale * rx[i];
            for (int r = 0; r < nr; ++r) {
                const double *g = Gr + (size_t)r * n; const int *zi = nzs + (size_t)r * n;
                double rc = -sl[r] * zl[r];
                if (pass == 1) rc += sigma * mu - dsa[r] * dza[r];
                dzl[r] = rc;
                double gg = rc / sl[r];
                for (int k = 0; k < nzc[r]; ++k) rhs[zi[k]] -= g[zi[k]] * gg;
            }
            { double wq[3]; nt_apply(&nt, tq, wq, 1);
              for (int i = 0; i < n; ++i) rhs[i] += Br[i] * wq[0] + Br[n + i] * wq[1] + Br[2 * n + i] * wq[2]; }
            memcpy(dv, rhs, sizeof(double) * n);
            chol_fwd(L, n, dv); chol_bwd(L, n, dv);
            double en_prev = 1e300;
            for (int ref = 0; ref < (getenv("GCSO_NOREF") ? 0 : 3); ++ref) { /* iterative refinement against H */
                double en = 0.0, bn = 0.0;
                for (int i = 0; i < n; ++i) {
                    double s = rhs[i];
                    for (int j = 0; j < n; ++j) s -= H[(size_t)i * n + j] * dv[j];
                    e1[i] = s; en = fmax(en, fabs(s)); bn = fmax(bn, fabs(rhs[i]));
                }
                if (en <= 1e-14 * (bn + 1e-300) || en >= 0.5 * en_prev) break;
                en_prev = en;
                memcpy(cv, e1, sizeof(double) * n);
                chol_fwd(L, n, cv); chol_bwd(L, n, cv);
                for (int i = 0; i < n; ++i) dv[i] += cv[i];
            }
            double tmax = 0.0;
            for (int r = 0; r < nr; ++r) {
                const double *g = Gr + (size_t)r * n; const int *zi = nzs + (size_t)r * n;
                double gd = 0.0;
                for (int k = 0; k < nzc[r]; ++k) gd += g[zi[k]] * dv[zi[k]];
                double ds = -gd, dz = (dzl[r] - zl[r] * ds) / sl[r];
                dsl[r] = ds; dzl[r] = dz;
                tmax = fmax(tmax, fmax(-ds / sl[r], -dz / zl[r]));
            }
            double bdp[3] = {0, 0, 0};
            for (int a = 0; a < 3; ++a) for (int c = 0; c < n; ++c) bdp[a] += Br[a * n + c] * dv[c];
            nt_apply(&nt, bdp, dsq_s, 1);
            for (int k = 0; k < 3; ++k) dzq_s[k] = tq[k] - dsq_s[k];
            tmax = fmax(tmax, fmax(soc_max_step(nt.lam, dsq_s), soc_max_step(nt.lam, dzq_s)));
            if (pass == 0) {
                double a = tmax <= 0.0 ? 1.0 : fmin(1.0, 1.0 / tmax);
                sigma = (1.0 - a) * (1.0 - a) * (1.0 - a);
                { /* centrality-aware floor on sigma (LOQO's rule): re-centre when a product lags */
                    double mn = sq[0] * zq[0] + sq[1] * zq[1] + sq[2] * zq[2];
                    for (int r = 0; r < nr; ++r) mn = fmin(mn, sl[r] * zl[r]);
                    double xi = fmax(mn / mu, 1e-300);
                    double c = fmin(0.05 * (1.0 - xi) / xi, 2.0);
                    sigma = fmax(sigma, loqo_c * c * c * c);
                }
                memcpy(dsa, dsl, sizeof(double) * nr); memcpy(dza, dzl, sizeof(double) * nr);
                memcpy(dsq_a, dsq_s, sizeof(dsq_s)); memcpy(dzq_a, dzq_s, sizeof(dzq_s));
            } else {
                alpha = tmax <= 0.0 ? 1.0 : fmin(1.0, 0.99 / tmax);
            }
        }
        double dzq[3], dsq[3];
        nt_apply(&nt, dzq_s, dzq, 1); nt_apply(&nt, dsq_s, dsq, 0);
        for (int bt = 0; bt < 30; ++bt) { /* neighbourhood safeguard */
            double sum = 0.0, mn = 1e300;
            for (int r = 0; r < nr; ++r) {
                double pr = (sl[r] + alpha * dsl[r]) * (zl[r] + alpha * dzl[r]);
                sum += pr; if (pr < mn) mn = pr;
            }
            double pq = 0.0;
            for (int k = 0; k < 3; ++k) pq += (sq[k] + alpha * dsq[k]) * (zq[k] + alpha * dzq[k]);
            sum += pq; if (pq < mn) mn = pq;
            if (mn >= gamma_nb * sum / deg) break;
            alpha *= 0.7;
        }
        if (g_verbose) printf("      sigma %.3e alpha %.3e lifted %d\n", sigma, alpha, lifted);
        { int finite = (alpha == alpha);
          for (int i = 0; i < n; ++i) if (!(dv[i] == dv[i]) || isinf(dv[i])) finite = 0;
          if (!finite) { info.status = 2; break; } }   /* keep the last good iterate */
        for (int i = 0; i < n; ++i) v[i] += alpha * dv[i];
        for (int r = 0; r < nr; ++r) zl[r] += alpha * dzl[r];
        for (int k = 0; k < 3; ++k) zq[k] += alpha * dzq[k];
    }
    if (info.status != 0) { memcpy(v, vbest, sizeof(double) * n); info.gap = best_gap; info.dres = best_dres; }
    free(vbest);
    for (int i = 0; i < nvar; ++i) {
        double s = ns.up[i];
        for (int c = 0; c < n; ++c) s += ns.N[(size_t)i * n + c] * v[c];
        u[i] = s;
    }
    free(Gr); free(hr); free(nzs); free(nzc); free(Pr); free(qr); free(Br); free(v);
    free(zl); free(sl); free(dsl); free(dzl); free(dsa); free(dza); free(H); free(L);
    free(rx); free(rhs); free(dv); free(e1); free(cv); free(ns.N); free(ns.up);
    return info;
}

/* ---------------------------------------------------------------- ADMM driver */
static void classify(Oracle *o) {
    int nV = o->nV;
    for (int v = 0; v < nV; ++v) {
        int h0 = o->he_off[v], h1 = o->he_off[v + 1];
        int is_s = (v == o->src), is_t = (v == o->dst);
        int live_in = 0, live_out = 0;
        for (int h = h0; h < h1; ++h) {
            int zero = (is_s && !o->he_out[h]) || (is_t && o->he_out[h]);
            o->he_zero[h] = (unsigned char)zero;
            if (!zero) { if (o->he_out[h]) live_out++; else live_in++; }
        }
        int dead = (!is_s && live_in == 0) || (!is_t && live_out == 0);
        if (dead) for (int h = h0; h < h1; ++h) o->he_zero[h] = 1;
        o->vtype[v] = dead ? 3 : (is_s ? 1 : (is_t ? 2 : 0));
    }
}

Oracle *gcso_create(int nV, int nE, const int *poly_off, const double *polyA, const double *polyb,
                    const int *he_off, const int *he_edge, const int *he_out,
                    const int *edge_he_tail, const int *edge_he_head, int src, int dst, const double *cent) {
    Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
    int H = 2 * nE, M = poly_off[nV];
    o->nV = nV; o->nE = nE; o->src = src; o->dst = dst;
#define DUP(dst_, src_, n_, T_) do { dst_ = (T_ *)malloc(sizeof(T_) * (size_t)((n_) > 0 ? (n_) : 1)); memcpy(dst_, src_, sizeof(T_) * (size_t)(n_)); } while (0)
    DUP(o->poly_off, poly_off, nV + 1, int); DUP(o->polyA, polyA, 2 * M, double); DUP(o->polyb, polyb, M, double);
    DUP(o->he_off, he_off, nV + 1, int); DUP(o->he_edge, he_edge, H, int); DUP(o->he_out, he_out, H, int);
    DUP(o->edge_he_tail, edge_he_tail, nE, int); DUP(o->edge_he_head, edge_he_head, nE, int);
    DUP(o->cent, cent, 2 * nV, double);
    o->rho0 = 1.0; o->tau_incr = 2.0; o->tau_decr = 2.0; o->nu = 10.0; o->frac = 0.1;
    o->eps_abs = 1e-4; o->eps_rel = 1e-3; o->max_it = 1000; o->inner_tol = 1e-9; o->inner_max_iter = 60;
    o->xc = (double *)calloc((size_t)5 * H + 1, sizeof(double)); o->mu = (double *)calloc((size_t)5 * H + 1, sizeof(double));
    o->z = (double *)calloc((size_t)5 * nE + 1, sizeof(double)); o->zprev = (double *)calloc((size_t)5 * nE + 1, sizeof(double));
    o->x_v = (double *)calloc((size_t)4 * nV, sizeof(double)); o->z_v = (double *)calloc((size_t)4 * nV, sizeof(double));
    o->y_v = (double *)calloc(nV, sizeof(double));
    o->he_zero = (unsigned char *)calloc(H + 1, 1); o->vtype = (unsigned char *)calloc(nV, 1);
    o->rho = o->rho0; o->it = 0; o->opt = 0;
    o->hist_cap = 1024;
    o->rho_seq = (double *)malloc(sizeof(double) * o->hist_cap); o->pri_seq = (double *)malloc(sizeof(double) * o->hist_cap);
    o->dual_seq = (double *)malloc(sizeof(double) * o->hist_cap);
    o->rho_seq[0] = o->rho; o->pri_seq[0] = 0.0; o->dual_seq[0] = 0.0; /* :637-639 */
    classify(o);
    return o;
}

void gcso_destroy(Oracle *o) {
    free(o->poly_off); free(o->polyA); free(o->polyb); free(o->he_off); free(o->he_edge); free(o->he_out);
    free(o->edge_he_tail); free(o->edge_he_head); free(o->cent); free(o->xc); free(o->mu); free(o->z); free(o->zprev);
    free(o->x_v); free(o->z_v); free(o->y_v); free(o->he_zero); free(o->vtype); free(o->rho_seq); free(o->pri_seq); free(o->dual_seq);
    free(o);
}

void gcso_set_params(Oracle *o, double rho0, double tau_incr, double tau_decr, double nu, double frac,
                     double eps_abs, double eps_rel, int max_it, double inner_tol, int inner_max_iter) {
    o->rho0 = rho0; o->tau_incr = tau_incr; o->tau_decr = tau_decr; o->nu = nu; o->frac = frac;
    o->eps_abs = eps_abs; o->eps_rel = eps_rel; o->max_it = max_it; o->inner_tol = inner_tol; o->inner_max_iter = inner_max_iter;
    if (o->it == 0) { o->rho = rho0; o->rho_seq[0] = rho0; }
}

/* x-update of one vertex (admm_solver_v3.py:352-466 + scatter :492-522) */
static int vertex_update(Oracle *o, int v, long *iters) {
    int h0 = o->he_off[v], h1 = o->he_off[v + 1];
    int m = o->poly_off[v + 1] - o->poly_off[v];
    const double *A = o->polyA + 2 * o->poly_off[v], *b = o->polyb + o->poly_off[v];
    int type = o->vtype[v];
    /* forced-zero half-edges */
    for (int h = h0; h < h1; ++h) if (o->he_zero[h]) {
        int e = o->he_edge[h];
        double *x = o->xc + 5 * h;
        x[0] = x[1] = x[2] = x[3] = x[4] = 0.0;
        if (!o->he_out[h]) { /* other endpoint's first point is free: sits on its target */
            x[0] = o->z[5 * e] + o->mu[5 * h]; x[1] = o->z[5 * e + 1] + o->mu[5 * h + 1];
        }
    }
    if (type == 3) {
        for (int k = 0; k < 4; ++k) { o->z_v[4 * v + k] = 0.0; o->x_v[4 * v + k] = o->cent[2 * v + (k & 1)]; }
        o->y_v[v] = 0.0;
        return 0;
    }
    int dmax = h1 - h0, d = 0;
    int *out = (int *)malloc(sizeof(int) * (dmax + 1)), *hid = (int *)malloc(sizeof(int) * (dmax + 1));
    double *tgt = (double *)malloc(sizeof(double) * 5 * (dmax + 1));
    for (int h = h0; h < h1; ++h) if (!o->he_zero[h]) {
        int e = o->he_edge[h];
        out[d] = o->he_out[h]; hid[d] = h;
        for (int k = 0; k < 5; ++k) tgt[5 * d + k] = o->z[5 * e + k] + o->mu[5 * h + k];
        d++;
    }
    VProg *p = vprog_build(m, A, b, d, out, type != 0, o->cent + 2 * v, o->rho, tgt);
    double *u = (double *)malloc(sizeof(double) * p->nvar);
    IpmInfo info = vprog_solve(p, out, o->inner_tol, o->inner_max_iter, u);
    *iters += info.iters;
    int bad = (info.status != 0) && !(info.dres <= 1e-6 && info.pres <= 1e-6 && info.gap <= 1e-6);
    if (bad && getenv("GCSO_DUMP")) {
        char path[512]; snprintf(path, sizeof path, "%s_v%d_it%d.bin", getenv("GCSO_DUMP"), v, o->it);
        FILE *fh = fopen(path, "wb");
        if (fh) {
            int hdr[4] = {m, d, type != 0, 0};
            fwrite(hdr, sizeof(int), 4, fh); fwrite(A, sizeof(double), 2 * m, fh); fwrite(b, sizeof(double), m, fh);
            fwrite(out, sizeof(int), d, fh); fwrite(o->cent + 2 * v, sizeof(double), 2, fh); fwrite(&o->rho, sizeof(double), 1, fh);
            fwrite(tgt, sizeof(double), 5 * d, fh); fclose(fh);
            fprintf(stderr, "oracle: inner solve failed v=%d it=%d status=%d gap=%.2e dres=%.2e pres=%.2e -> %s\n", v, o->it, info.status, info.gap, info.dres, info.pres, path);
        }
    }
    for (int k = 0; k < 4; ++k) { o->x_v[4 * v + k] = u[p->X + k]; o->z_v[4 * v + k] = u[p->Z + k]; }
    o->y_v[v] = type ? 1.0 : u[p->YV];
    for (int j = 0; j < d; ++j) {
        const double *w = u + p->nc + 5 * j;
        double *x = o->xc + 5 * hid[j];
        if (out[j]) { x[0] = w[0]; x[1] = w[1]; x[2] = w[2]; x[3] = w[3]; }
        else { x[0] = tgt[5 * j]; x[1] = tgt[5 * j + 1]; x[2] = w[0]; x[3] = w[1]; }
        x[4] = w[4];
    }
    free(u); vprog_free(p); free(out); free(hid); free(tgt);
    return bad;
}

/* k passes of the loop body :655-733; stops early on convergence if check_stop. returns iterations done */
int gcso_step(Oracle *o, int k, int check_stop) {
    int nV = o->nV, nE = o->nE, H = 2 * nE, done = 0;
    for (int s = 0; s < k; ++s) {
        o->it++;
        int it = o->it;
        long iters = 0; int fails = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : iters, fails)
        for (int v = 0; v < nV; ++v) fails += vertex_update(o, v, &iters);
        o->inner_iters += iters; o->inner_fail += fails;
        memcpy(o->zprev, o->z, sizeof(double) * 5 * nE);
        double dz2 = 0.0, z2 = 0.0;
        for (int e = 0; e < nE; ++e) { /* :543-562 */
            const double *a = o->xc + 5 * o->edge_he_tail[e], *b = o->xc + 5 * o->edge_he_head[e];
            for (int c = 0; c < 5; ++c) {
                double zn = 0.5 * (a[c] + b[c]);
                double dd = zn - o->zprev[5 * e + c];
                o->z[5 * e + c] = zn; dz2 += dd * dd; z2 += zn * zn;
            }
        }
        double r2 = 0.0, x2 = 0.0;
        for (int h = 0; h < H; ++h) { /* :594, :598 */
            int e = o->he_edge[h];
            for (int c = 0; c < 5; ++c) {
                double r = o->z[5 * e + c] - o->xc[5 * h + c];
                o->mu[5 * h + c] += r; r2 += r * r; x2 += o->xc[5 * h + c] * o->xc[5 * h + c];
            }
        }
        double pri = sqrt(r2), dual = o->rho * sqrt(2.0 * dz2); /* :598, :602 */
        double scale = 1.0;
        if (pri >= o->nu * dual && it < o->frac * o->max_it) { o->rho *= o->tau_incr; scale = 1.0 / o->tau_incr; } /* :703-705 */
        else if (dual >= o->nu * pri && it < o->frac * o->max_it) { o->rho *= 1.0 / o->tau_decr; scale = o->tau_incr; } /* :706-708 */
        double m2 = 0.0;
        for (int i = 0; i < 5 * H; ++i) { o->mu[i] *= scale; m2 += o->mu[i] * o->mu[i]; }
        if (it + 1 >= o->hist_cap) {
            o->hist_cap *= 2;
            o->rho_seq = (double *)realloc(o->rho_seq, sizeof(double) * o->hist_cap);
            o->pri_seq = (double *)realloc(o->pri_seq, sizeof(double) * o->hist_cap);
            o->dual_seq = (double *)realloc(o->dual_seq, sizeof(double) * o->hist_cap);
        }
        o->rho_seq[it] = o->rho; o->pri_seq[it] = pri; o->dual_seq[it] = dual;
        double nAx = sqrt(x2), nBz = sqrt(2.0 * z2);
        double eps_pri = sqrt((double)(9 * nV + 18 * nE)) * o->eps_abs + o->eps_rel * fmax(nAx, nBz); /* :605-610 */
        double eps_dual = sqrt((double)(10 * nE)) * o->eps_abs + o->eps_rel * sqrt(m2);                /* :613-614 */
        done++;
        if (!(pri == pri) || !(dual == dual) || isinf(pri) || isinf(dual)) { o->diverged = 1; break; }
        if (pri < eps_pri && dual < eps_dual) { o->opt = 1; if (check_stop) break; } /* :712 */
    }
    return done;
}

void gcso_get_info(const Oracle *o, int *it, int *opt, double *rho, long *inner_iters, int *inner_fail) {
    *it = o->it; *opt = o->opt; *rho = o->rho; *inner_iters = o->inner_iters; *inner_fail = o->inner_fail;
}
void gcso_get_history(const Oracle *o, double *rho, double *pri, double *dual) {
    memcpy(rho, o->rho_seq, sizeof(double) * (o->it + 1)); memcpy(pri, o->pri_seq, sizeof(double) * (o->it + 1));
    memcpy(dual, o->dual_seq, sizeof(double) * (o->it + 1));
}
void gcso_get_state(const Oracle *o, double *xc, double *mu, double *z) {
    memcpy(xc, o->xc, sizeof(double) * 10 * o->nE); memcpy(mu, o->mu, sizeof(double) * 10 * o->nE);
    memcpy(z, o->z, sizeof(double) * 5 * o->nE);
}
void gcso_set_state(Oracle *o, const double *xc, const double *mu, const double *z, double rho, int it) {
    memcpy(o->xc, xc, sizeof(double) * 10 * o->nE); memcpy(o->mu, mu, sizeof(double) * 10 * o->nE);
    memcpy(o->z, z, sizeof(double) * 5 * o->nE); o->rho = rho; o->it = it;
}
void gcso_get_solution(const Oracle *o, double *x_v, double *z_v, double *y_v) {
    memcpy(x_v, o->x_v, sizeof(double) * 4 * o->nV); memcpy(z_v, o->z_v, sizeof(double) * 4 * o->nV);
    memcpy(y_v, o->y_v, sizeof(double) * o->nV);
}
/* x-update only (for per-kernel parity tests) */
int gcso_vertex_update_all(Oracle *o) {
    long iters = 0; int fails = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : iters, fails)
    for (int v = 0; v < o->nV; ++v) fails += vertex_update(o, v, &iters);
    o->inner_iters += iters; o->inner_fail += fails;
    return fails;
}
double gcso_cost(const Oracle *o) { /* GCS_utils.py:184-211 */
    double c = 0.0;
    for (int v = 0; v < o->nV; ++v) c += hypot(o->z_v[4 * v] - o->z_v[4 * v + 2], o->z_v[4 * v + 1] - o->z_v[4 * v + 3]);
    for (int e = 0; e < o->nE; ++e) c += EDGE_PENALTY * o->z[5 * e + 4];
    return c;
}
void gcso_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int gcso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* debugging: solve one dumped vertex program */
int gcso_solve_raw(int m, const double *A, const double *b, int d, const int *out, int term, const double *cent,
                   double rho, const double *tgt, double tol, int max_iter, double *u_out, double *info_out) {
    VProg *p = vprog_build(m, A, b, d, out, term, cent, rho, tgt);
    double *u = (double *)malloc(sizeof(double) * p->nvar);
    IpmInfo info = vprog_solve(p, out, tol, max_iter, u);
    memcpy(u_out, u, sizeof(double) * p->nvar);
    info_out[0] = info.iters; info_out[1] = info.status; info_out[2] = info.gap; info_out[3] = info.dres; info_out[4] = info.pres;
    int nvar = p->nvar;
    free(u); vprog_free(p);
    return nvar;
}
