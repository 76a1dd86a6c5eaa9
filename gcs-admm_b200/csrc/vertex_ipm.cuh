// Per-vertex prox solve of the full-vertex-split ADMM — one WARP per vertex, state in shared memory.
//
// Solves reference admm_solver_v3.py:352-466 (the program handed to MOSEK at :490) for one
// vertex in its reduced, equality-free form (DESIGN.md "K1"):
//   u-space  : x(4) t z(4) y_v | (a1 a2 y) per live half-edge            (structured rows live here)
//   v-space  : x(4) t | (a1 a2 y) per live half-edge except one dependent edge j*   (dense Newton system)
//   u = N v + up  eliminates C6/C7 (:450-464):  (z, y_v) = sum over the secondary group,
//   w_{j*} = (z, y_v) - sum over the other primary edges.  For 's'/'t' the secondary group is the
//   virtual edge (x, 1).
// Method: primal-feasible start, Mehrotra predictor-corrector with a centrality floor on sigma and a
// wide-neighbourhood step safeguard, Nesterov-Todd scaling for the single second-order cone
// (t >= |z1 - z2|), dense Cholesky with pivot lifting.  Everything is fp64.
//
// The file compiles two ways:
//   * nvcc (device): GCS_LANE_LOOP strides the 32 lanes of a warp, reductions are shuffles;
//   * g++  with GCS_EMULATE (tests only): a "warp" is one host thread, lane loops run serially and
//     reductions are identities.  Phases separated by GCS_SYNC() never carry per-lane state, so both
//     builds execute the same arithmetic.  The emulation build is a debugging aid for the CPU test
//     suite; the product library has no CPU path.
#pragma once
#include <math.h>
#ifdef GCS_EMULATE
#include <stdio.h>
#include <stdlib.h>
#endif

#ifdef GCS_EMULATE
#define GCS_DEV static inline
#define GCS_LANE_LOOP(i, n) for (int i = 0; i < (n); ++i)
#define GCS_SYNC() ((void)0)
static inline double gcs_warp_sum(double x) { return x; }
static inline double gcs_warp_max(double x) { return x; }
static inline double gcs_warp_min(double x) { return x; }
#else
#define GCS_DEV __device__ __forceinline__
#define GCS_LANE_LOOP(i, n) for (int i = lane; i < (n); i += 32)
#define GCS_SYNC() __syncwarp()
__device__ __forceinline__ double gcs_warp_sum(double x) {
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double gcs_warp_max(double x) {
#pragma unroll
    for (int o = 16; o; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ double gcs_warp_min(double x) {
#pragma unroll
    for (int o = 16; o; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
#endif

#define GCS_EDGE_PENALTY 1e-4  // reference admm_solver_v3.py:388
#define GCS_NCORE 10           // u-space core: x(4) t z(4) y_v
#define GCS_UX 0
#define GCS_UT 4
#define GCS_UZ 5
#define GCS_UYV 9
#define GCS_EPP 16             // doubles per edge-point partial record
#define GCS_NB_GAMMA 1e-5      // width of the central-path neighbourhood
#define GCS_LOQO_C 0.02        // weight of the centrality-aware floor on sigma

// Scratch layout (offsets in doubles) of one warp, for <= dcap live half-edges and <= mcap polytope rows.
struct GcsScratchLayout {
    int dcap, mcap, ncap, nucap, ldh;
    int H, M, B, C, u, dua, du, gu, ru, Pu, qu, v, dv, rv, vbest, zr, dsr, dzr, zc, dsc, dzc, zy, dsy, dzy, sy;
    int ep, cp, A, b, AA, tgt, ints, total;
};

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline GcsScratchLayout gcs_scratch_layout(int dcap, int mcap) {
    GcsScratchLayout L;
    if (dcap < 1) dcap = 1;
    if (mcap < 1) mcap = 1;
    L.dcap = dcap; L.mcap = mcap;
    L.ncap = 5 * dcap;            // v-space: 5 + 5 (d - 1)
    L.nucap = GCS_NCORE + 5 * dcap;
    L.ldh = L.ncap | 1;           // odd row stride: fewer shared-memory bank conflicts
    int o = 0;
    L.H = o; o += L.ncap * L.ldh;
    L.M = o; o += 25 * dcap;
    L.B = o; o += 20 * dcap;
    L.C = o; o += 100;
    L.u = o; o += L.nucap;  L.dua = o; o += L.nucap;  L.du = o; o += L.nucap;
    L.gu = o; o += L.nucap; L.ru = o; o += L.nucap;  L.Pu = o; o += L.nucap;  L.qu = o; o += L.nucap;
    L.v = o; o += L.ncap;   L.dv = o; o += L.ncap;   L.rv = o; o += L.ncap;   L.vbest = o; o += L.ncap;
    int nr = dcap * 2 * mcap * 2;
    L.zr = o; o += nr;  L.dsr = o; o += nr;  L.dzr = o; o += nr;
    int ncr = 2 * mcap * 2;
    L.zc = o; o += ncr; L.dsc = o; o += ncr; L.dzc = o; o += ncr;
    L.zy = o; o += dcap + 1; L.dsy = o; o += dcap + 1; L.dzy = o; o += dcap + 1; L.sy = o; o += dcap + 1;
    L.ep = o; o += GCS_EPP * 2 * dcap;
    L.cp = o; o += 16;
    L.A = o; o += 2 * mcap; L.b = o; o += mcap; L.AA = o; o += 3 * mcap;
    L.tgt = o; o += 5 * dcap;
    L.ints = o; o += (3 * dcap + 1) / 2 + 1;   // int arrays out / prim / hid packed behind the doubles
    L.total = o;
    return L;
}

struct GcsVertexIn {
    int m;                 // polytope rows (A, b already staged in scratch)
    int d;                 // live half-edges (out / targets already staged in scratch)
    int type;              // 0 generic, 1 source, 2 target
    double cx, cy;         // strictly interior point of the polytope
    double rho;
    double tol; int max_iter;
};

struct GcsVertexOut { int iters; int status; double gap, dres; };

struct GcsNT { double w0, w1, w2, beta, l0, l1, l2; };

GCS_DEV double gcs_jnorm2(double s0, double s1, double s2) { double n1 = hypot(s1, s2); return (s0 - n1) * (s0 + n1); }
GCS_DEV void gcs_nt_apply(const GcsNT &S, double x0, double x1, double x2, bool inverse, double &y0, double &y1, double &y2) {
    double w1 = inverse ? -S.w1 : S.w1, w2 = inverse ? -S.w2 : S.w2;
    double t = w1 * x1 + w2 * x2, c = x0 + t / (1.0 + S.w0);
    double sc = inverse ? 1.0 / S.beta : S.beta;
    y0 = (S.w0 * x0 + t) * sc; y1 = (x1 + c * w1) * sc; y2 = (x2 + c * w2) * sc;
}
GCS_DEV void gcs_nt_build(GcsNT &S, const double *s, const double *z) {
    double sn = sqrt(fmax(gcs_jnorm2(s[0], s[1], s[2]), 1e-300)), zn = sqrt(fmax(gcs_jnorm2(z[0], z[1], z[2]), 1e-300));
    double sb0 = s[0] / sn, sb1 = s[1] / sn, sb2 = s[2] / sn, zb0 = z[0] / zn, zb1 = z[1] / zn, zb2 = z[2] / zn;
    double gamma = sqrt(0.5 * (1.0 + sb0 * zb0 + sb1 * zb1 + sb2 * zb2));
    S.w0 = (sb0 + zb0) / (2 * gamma); S.w1 = (sb1 - zb1) / (2 * gamma); S.w2 = (sb2 - zb2) / (2 * gamma);
    S.beta = sqrt(sn / zn);
    gcs_nt_apply(S, z[0], z[1], z[2], false, S.l0, S.l1, S.l2);
}
GCS_DEV double gcs_soc_max_step(const GcsNT &S, const double *d) {
    double nrm = sqrt(fmax(gcs_jnorm2(S.l0, S.l1, S.l2), 1e-300));
    double l0 = S.l0 / nrm, l1 = S.l1 / nrm, l2 = S.l2 / nrm;
    double c0 = l0 * d[0] - l1 * d[1] - l2 * d[2], f = (c0 + d[0]) / (l0 + 1.0);
    return (hypot(d[1] - f * l1, d[2] - f * l2) - c0) / nrm;
}
GCS_DEV void gcs_soc_div(const GcsNT &S, const double *v, double *x) {  // lam o x = v
    double det = gcs_jnorm2(S.l0, S.l1, S.l2);
    double x0 = (S.l0 * v[0] - S.l1 * v[1] - S.l2 * v[2]) / det;
    x[0] = x0; x[1] = (v[1] - x0 * S.l1) / S.l0; x[2] = (v[2] - x0 * S.l2) / S.l0;
}

GCS_DEV int gcs_uw(int j) { return GCS_NCORE + 5 * j; }   // u-space offset of live half-edge j

// ---- the null-space map ---------------------------------------------------------------------
// forward:  u = N v (+ up when `affine`)
GCS_DEV void gcs_forward(const double *v, double *u, int d, int jstar, const int *prim, bool term, bool affine, int lane) {
    GCS_LANE_LOOP(c, 5) {
        if (c < 4) u[GCS_UX + c] = v[c]; else u[GCS_UT] = v[4];
    }
    GCS_LANE_LOOP(q, 5 * (d - 1)) {
        int jj = q / 5, c = q - 5 * jj, j = jj < jstar ? jj : jj + 1;
        u[gcs_uw(j) + c] = v[5 + q];
    }
    GCS_LANE_LOOP(c, 5) {      // (z, y_v) and the dependent edge
        double zy = term ? (c < 4 ? v[c] : (affine ? 1.0 : 0.0)) : 0.0, other = 0.0;
        for (int jj = 0; jj < d - 1; ++jj) {
            int j = jj < jstar ? jj : jj + 1;
            double w = v[5 + 5 * jj + c];
            if (prim[j]) other += w; else zy += w;
        }
        u[GCS_UZ + c] = zy;     // GCS_UZ + 4 == GCS_UYV
        u[gcs_uw(jstar) + c] = zy - other;
    }
    GCS_SYNC();
}
// adjoint:  out = N' g
GCS_DEV void gcs_adjoint(const double *g, double *out, int d, int jstar, const int *prim, bool term, int lane) {
    GCS_LANE_LOOP(c, 5) {
        if (c < 4) out[c] = g[GCS_UX + c] + (term ? g[GCS_UZ + c] + g[gcs_uw(jstar) + c] : 0.0); else out[4] = g[GCS_UT];
    }
    GCS_LANE_LOOP(q, 5 * (d - 1)) {
        int jj = q / 5, c = q - 5 * jj, j = jj < jstar ? jj : jj + 1;
        double gs = g[gcs_uw(jstar) + c];
        out[5 + q] = g[gcs_uw(j) + c] + (prim[j] ? -gs : g[GCS_UZ + c] + gs);
    }
    GCS_SYNC();
}

// H_u(a, b) from its structured pieces (a, b are u-space indices)
GCS_DEV double gcs_hu(const double *M, const double *B, const double *C, int a, int b) {
    if (a < GCS_NCORE && b < GCS_NCORE) return C[a * 10 + b];
    if (a >= GCS_NCORE && b >= GCS_NCORE) {
        int ja = (a - GCS_NCORE) / 5, jb = (b - GCS_NCORE) / 5;
        if (ja != jb) return 0.0;
        return M[25 * ja + 5 * (a - GCS_NCORE - 5 * ja) + (b - GCS_NCORE - 5 * jb)];
    }
    if (a > b) { int t = a; a = b; b = t; }   // a core, b block
    if (a >= 4) return 0.0;                   // only x couples with the edge blocks (C4)
    int jb = (b - GCS_NCORE) / 5;
    return B[20 * jb + 5 * a + (b - GCS_NCORE - 5 * jb)];
}
// column p of N as up to three (u-index, coefficient) pairs; returns the count
GCS_DEV int gcs_ncol(int p, int jstar, const int *prim, bool term, int *idx, double *cf) {
    if (p < 4) {
        idx[0] = GCS_UX + p; cf[0] = 1.0;
        if (!term) return 1;
        idx[1] = GCS_UZ + p; cf[1] = 1.0; idx[2] = gcs_uw(jstar) + p; cf[2] = 1.0;
        return 3;
    }
    if (p == 4) { idx[0] = GCS_UT; cf[0] = 1.0; return 1; }
    int jj = (p - 5) / 5, c = p - 5 - 5 * jj, j = jj < jstar ? jj : jj + 1;
    idx[0] = gcs_uw(j) + c; cf[0] = 1.0;
    idx[1] = gcs_uw(jstar) + c;
    if (prim[j]) { cf[1] = -1.0; return 2; }
    cf[1] = 1.0; idx[2] = GCS_UZ + c; cf[2] = 1.0;
    return 3;
}

// ---- one pass over the inequality rows ---------------------------------------------------------
//   mode 0: Hessian pieces, gradient G'z into gout, gap (r1) and smallest complementarity product (r0)
//   mode 1: predictor step ratio for direction dua (r0 = max(-ds/s, -dz/z))
//   mode 2: corrector right-hand side  -G'(rc/s)  into gout        (rc uses dua and sigmu)
//   mode 3: corrector direction du: store ds, dz per row; r0 = max ratio
//   mode 4: neighbourhood statistics for step alpha: r0 = min, r1 = sum of (s + a ds)(z + a dz)
//   mode 5: z += alpha dz
struct GcsRowsArgs { int mode; double sigmu, alpha; double *gout; };

GCS_DEV void gcs_rows(const GcsScratchLayout &L, double *S, int m, int d, bool term, const GcsRowsArgs &ar,
                      double &r0, double &r1, int lane) {
    const double *A = S + L.A, *b = S + L.b, *AA = S + L.AA;
    const double *u = S + L.u, *du = S + L.du, *dua = S + L.dua;
    double *gout = ar.gout;
    double acc_max = 0.0, acc_min = 1e300, acc_sum = 0.0;
    const int mode = ar.mode;
    const bool need_p = (mode == 1 || mode == 2 || mode == 3), need_d = (mode == 3);
    // ---- edge-point items (j, i): rows C3 (:434-436) and C4 (:438-440) of half-edge j, point i
    GCS_LANE_LOOP(e, 2 * d) {
        const int j = e >> 1, i = e & 1;
        const int ao = gcs_uw(j) + 2 * i, yo = gcs_uw(j) + 4, xo = GCS_UX + 2 * i;
        const double a0 = u[ao], a1 = u[ao + 1], y = u[yo], x0 = u[xo], x1 = u[xo + 1];
        double da0 = 0, da1 = 0, dy = 0, dx0 = 0, dx1 = 0, pa0 = 0, pa1 = 0, py = 0, px0 = 0, px1 = 0;
        if (need_d) { da0 = du[ao]; da1 = du[ao + 1]; dy = du[yo]; dx0 = du[xo]; dx1 = du[xo + 1]; }
        if (need_p) { pa0 = dua[ao]; pa1 = dua[ao + 1]; py = dua[yo]; px0 = dua[xo]; px1 = dua[xo + 1]; }
        double *zr = S + L.zr + e * m * 2, *dsr = S + L.dsr + e * m * 2, *dzr = S + L.dzr + e * m * 2;
        double Maa0 = 0, Maa1 = 0, Maa2 = 0, May0 = 0, May1 = 0, Myy = 0, Bxa0 = 0, Bxa1 = 0, Bxa2 = 0, Bxy0 = 0, Bxy1 = 0;
        double ga0 = 0, ga1 = 0, gy = 0, gx0 = 0, gx1 = 0;
        for (int k = 0; k < m; ++k) {
            const double A0 = A[2 * k], A1 = A[2 * k + 1], bk = b[k];
            const double Aa = A0 * a0 + A1 * a1;
            const double s3 = y * bk - Aa, z3 = zr[2 * k];
            double s4 = 1.0, z4 = 0.0;
            if (!term) { s4 = (1.0 - y) * bk - (A0 * x0 + A1 * x1) + Aa; z4 = zr[2 * k + 1]; }
            if (mode == 0) {
                const double D3 = z3 / s3, D4 = z4 / s4, Ds = D3 + D4;
                Maa0 += Ds * AA[3 * k]; Maa1 += Ds * AA[3 * k + 1]; Maa2 += Ds * AA[3 * k + 2];
                May0 -= Ds * bk * A0; May1 -= Ds * bk * A1; Myy += Ds * bk * bk;
                Bxa0 -= D4 * AA[3 * k]; Bxa1 -= D4 * AA[3 * k + 1]; Bxa2 -= D4 * AA[3 * k + 2];
                Bxy0 += D4 * bk * A0; Bxy1 += D4 * bk * A1;
                const double zd = z3 - z4;
                ga0 += A0 * zd; ga1 += A1 * zd; gy -= bk * zd; gx0 += A0 * z4; gx1 += A1 * z4;
                const double p3 = s3 * z3;
                acc_sum += p3; acc_min = fmin(acc_min, p3);
                if (!term) { const double p4 = s4 * z4; acc_sum += p4; acc_min = fmin(acc_min, p4); }
            } else if (need_p) {
                // predictor quantities (rc = -s z):  ds = -g.dua,  dz = -z - z ds / s
                const double pAa = A0 * pa0 + A1 * pa1;
                const double ps3 = -(pAa - bk * py), ps4 = -((A0 * px0 + A1 * px1) - pAa + bk * py);
                const double pz3 = -z3 - z3 * ps3 / s3, pz4 = -z4 - z4 * ps4 / s4;
                if (mode == 1) {
                    acc_max = fmax(acc_max, fmax(-ps3 / s3, -pz3 / z3));
                    if (!term) acc_max = fmax(acc_max, fmax(-ps4 / s4, -pz4 / z4));
                } else {
                    const double rc3 = -s3 * z3 + ar.sigmu - ps3 * pz3, rc4 = -s4 * z4 + ar.sigmu - ps4 * pz4;
                    if (mode == 2) {
                        const double g3 = rc3 / s3, g4 = term ? 0.0 : rc4 / s4, gd = g3 - g4;
                        ga0 -= A0 * gd; ga1 -= A1 * gd; gy += bk * gd; gx0 -= A0 * g4; gx1 -= A1 * g4;
                    } else {
                        const double dAa = A0 * da0 + A1 * da1;
                        const double ds3 = -(dAa - bk * dy), dz3 = (rc3 - z3 * ds3) / s3;
                        dsr[2 * k] = ds3; dzr[2 * k] = dz3;
                        acc_max = fmax(acc_max, fmax(-ds3 / s3, -dz3 / z3));
                        if (!term) {
                            const double ds4 = -((A0 * dx0 + A1 * dx1) - dAa + bk * dy), dz4 = (rc4 - z4 * ds4) / s4;
                            dsr[2 * k + 1] = ds4; dzr[2 * k + 1] = dz4;
                            acc_max = fmax(acc_max, fmax(-ds4 / s4, -dz4 / z4));
                        }
                    }
                }
            } else if (mode == 4) {
                const double p3 = (s3 + ar.alpha * dsr[2 * k]) * (z3 + ar.alpha * dzr[2 * k]);
                acc_sum += p3; acc_min = fmin(acc_min, p3);
                if (!term) {
                    const double p4 = (s4 + ar.alpha * dsr[2 * k + 1]) * (z4 + ar.alpha * dzr[2 * k + 1]);
                    acc_sum += p4; acc_min = fmin(acc_min, p4);
                }
            } else {
                zr[2 * k] = z3 + ar.alpha * dzr[2 * k];
                if (!term) zr[2 * k + 1] = z4 + ar.alpha * dzr[2 * k + 1];
            }
        }
        if (mode == 0 || mode == 2) {
            double *ep = S + L.ep + GCS_EPP * e;
            if (mode == 0) {
                ep[0] = Maa0; ep[1] = Maa1; ep[2] = Maa2; ep[3] = May0; ep[4] = May1; ep[5] = Myy;
                ep[6] = Bxa0; ep[7] = Bxa1; ep[8] = Bxa2; ep[9] = Bxy0; ep[10] = Bxy1;
            }
            ep[11] = gy; ep[12] = gx0; ep[13] = gx1;
            gout[ao] = ga0; gout[ao + 1] = ga1;          // exclusive slots
        }
    }
    // ---- core items i: rows C1 (:420-422) and C2 (:424-426)
    GCS_LANE_LOOP(i, 2) {
        const int zo = GCS_UZ + 2 * i, xo = GCS_UX + 2 * i;
        const double z0 = u[zo], z1 = u[zo + 1], yv = u[GCS_UYV], x0 = u[xo], x1 = u[xo + 1];
        double dz0 = 0, dz1 = 0, dyv = 0, dx0 = 0, dx1 = 0, pz0 = 0, pz1 = 0, pyv = 0, px0 = 0, px1 = 0;
        if (need_d) { dz0 = du[zo]; dz1 = du[zo + 1]; dyv = du[GCS_UYV]; dx0 = du[xo]; dx1 = du[xo + 1]; }
        if (need_p) { pz0 = dua[zo]; pz1 = dua[zo + 1]; pyv = dua[GCS_UYV]; px0 = dua[xo]; px1 = dua[xo + 1]; }
        double *zc = S + L.zc + i * m * 2, *dsc = S + L.dsc + i * m * 2, *dzc = S + L.dzc + i * m * 2;
        double Zz0 = 0, Zz1 = 0, Zz2 = 0, Zy0 = 0, Zy1 = 0, Yy = 0, Xx0 = 0, Xx1 = 0, Xx2 = 0, Xy0 = 0, Xy1 = 0;
        double gz0 = 0, gz1 = 0, gyv = 0, gx0 = 0, gx1 = 0;
        for (int k = 0; k < m; ++k) {
            const double A0 = A[2 * k], A1 = A[2 * k + 1], bk = b[k];
            const double Az = A0 * z0 + A1 * z1;
            const double s1 = yv * bk - Az, c1 = zc[2 * k];
            double s2 = 1.0, c2 = 0.0;
            if (!term) { s2 = (1.0 - yv) * bk - (A0 * x0 + A1 * x1) + Az; c2 = zc[2 * k + 1]; }
            if (mode == 0) {
                const double D1 = c1 / s1, D2 = c2 / s2, Ds = D1 + D2;
                Zz0 += Ds * AA[3 * k]; Zz1 += Ds * AA[3 * k + 1]; Zz2 += Ds * AA[3 * k + 2];
                Zy0 -= Ds * bk * A0; Zy1 -= Ds * bk * A1; Yy += Ds * bk * bk;
                Xx0 += D2 * AA[3 * k]; Xx1 += D2 * AA[3 * k + 1]; Xx2 += D2 * AA[3 * k + 2];
                Xy0 += D2 * bk * A0; Xy1 += D2 * bk * A1;
                const double zd = c1 - c2;
                gz0 += A0 * zd; gz1 += A1 * zd; gyv -= bk * zd; gx0 += A0 * c2; gx1 += A1 * c2;
                const double p1 = s1 * c1;
                acc_sum += p1; acc_min = fmin(acc_min, p1);
                if (!term) { const double p2 = s2 * c2; acc_sum += p2; acc_min = fmin(acc_min, p2); }
            } else if (need_p) {
                const double pAz = A0 * pz0 + A1 * pz1;
                const double ps1 = -(pAz - bk * pyv), ps2 = -((A0 * px0 + A1 * px1) - pAz + bk * pyv);
                const double q1 = -c1 - c1 * ps1 / s1, q2 = -c2 - c2 * ps2 / s2;
                if (mode == 1) {
                    acc_max = fmax(acc_max, fmax(-ps1 / s1, -q1 / c1));
                    if (!term) acc_max = fmax(acc_max, fmax(-ps2 / s2, -q2 / c2));
                } else {
                    const double rc1 = -s1 * c1 + ar.sigmu - ps1 * q1, rc2 = -s2 * c2 + ar.sigmu - ps2 * q2;
                    if (mode == 2) {
                        const double g1 = rc1 / s1, g2 = term ? 0.0 : rc2 / s2, gd = g1 - g2;
                        gz0 -= A0 * gd; gz1 -= A1 * gd; gyv += bk * gd; gx0 -= A0 * g2; gx1 -= A1 * g2;
                    } else {
                        const double dAz = A0 * dz0 + A1 * dz1;
                        const double ds1 = -(dAz - bk * dyv), dd1 = (rc1 - c1 * ds1) / s1;
                        dsc[2 * k] = ds1; dzc[2 * k] = dd1;
                        acc_max = fmax(acc_max, fmax(-ds1 / s1, -dd1 / c1));
                        if (!term) {
                            const double ds2 = -((A0 * dx0 + A1 * dx1) - dAz + bk * dyv), dd2 = (rc2 - c2 * ds2) / s2;
                            dsc[2 * k + 1] = ds2; dzc[2 * k + 1] = dd2;
                            acc_max = fmax(acc_max, fmax(-ds2 / s2, -dd2 / c2));
                        }
                    }
                }
            } else if (mode == 4) {
                const double p1 = (s1 + ar.alpha * dsc[2 * k]) * (c1 + ar.alpha * dzc[2 * k]);
                acc_sum += p1; acc_min = fmin(acc_min, p1);
                if (!term) {
                    const double p2 = (s2 + ar.alpha * dsc[2 * k + 1]) * (c2 + ar.alpha * dzc[2 * k + 1]);
                    acc_sum += p2; acc_min = fmin(acc_min, p2);
                }
            } else {
                zc[2 * k] = c1 + ar.alpha * dzc[2 * k];
                if (!term) zc[2 * k + 1] = c2 + ar.alpha * dzc[2 * k + 1];
            }
        }
        if (mode == 0) {   // point-exclusive entries of C (zero-filled by the caller)
            double *c0 = S + L.C;
            const int X = GCS_UX + 2 * i, Z = GCS_UZ + 2 * i, Y = GCS_UYV;
            c0[Z * 10 + Z] = Zz0; c0[Z * 10 + Z + 1] = Zz1; c0[(Z + 1) * 10 + Z] = Zz1; c0[(Z + 1) * 10 + Z + 1] = Zz2;
            c0[Z * 10 + Y] = Zy0; c0[Y * 10 + Z] = Zy0; c0[(Z + 1) * 10 + Y] = Zy1; c0[Y * 10 + Z + 1] = Zy1;
            c0[X * 10 + X] = Xx0; c0[X * 10 + X + 1] = Xx1; c0[(X + 1) * 10 + X] = Xx1; c0[(X + 1) * 10 + X + 1] = Xx2;
            c0[X * 10 + Y] = Xy0; c0[Y * 10 + X] = Xy0; c0[(X + 1) * 10 + Y] = Xy1; c0[Y * 10 + X + 1] = Xy1;
            c0[X * 10 + Z] = -Xx0; c0[X * 10 + Z + 1] = -Xx1; c0[(X + 1) * 10 + Z] = -Xx1; c0[(X + 1) * 10 + Z + 1] = -Xx2;
            c0[Z * 10 + X] = -Xx0; c0[(Z + 1) * 10 + X] = -Xx1; c0[Z * 10 + X + 1] = -Xx1; c0[(Z + 1) * 10 + X + 1] = -Xx2;
        }
        if (mode == 0 || mode == 2) {
            gout[zo] = gz0; gout[zo + 1] = gz1;           // exclusive
            double *cp = S + L.cp + 8 * i;
            cp[0] = Yy; cp[1] = gyv; cp[2] = gx0; cp[3] = gx1;
        }
    }
    // ---- singles: y_e >= 0 (:377) and y_v <= 1 (:366)
    GCS_LANE_LOOP(j, d + 1) {
        if (j == d && term) continue;
        const int yo = j < d ? gcs_uw(j) + 4 : GCS_UYV;
        const double sgn = j < d ? 1.0 : -1.0;           // s = y   or   s = 1 - y_v ;  row g = -sgn on y
        const double s = j < d ? u[yo] : 1.0 - u[yo];
        double *zy = S + L.zy, *dsy = S + L.dsy, *dzy = S + L.dzy, *sy = S + L.sy;
        const double z = zy[j];
        if (mode == 0) {
            const double p = s * z;
            acc_sum += p; acc_min = fmin(acc_min, p);
            dsy[j] = z / s;          // D, consumed by the assembly
            sy[j] = -sgn * z;        // G'z contribution
        } else if (need_p) {
            const double ps = sgn * dua[yo], pz = -z - z * ps / s;
            if (mode == 1) acc_max = fmax(acc_max, fmax(-ps / s, -pz / z));
            else {
                const double rc = -s * z + ar.sigmu - ps * pz;
                if (mode == 2) sy[j] = sgn * rc / s;     // -G'(rc/s)
                else {
                    const double ds = sgn * du[yo], dz = (rc - z * ds) / s;
                    dsy[j] = ds; dzy[j] = dz;
                    acc_max = fmax(acc_max, fmax(-ds / s, -dz / z));
                }
            }
        } else if (mode == 4) {
            const double p = (s + ar.alpha * dsy[j]) * (z + ar.alpha * dzy[j]);
            acc_sum += p; acc_min = fmin(acc_min, p);
        } else {
            zy[j] = z + ar.alpha * dzy[j];
        }
    }
    GCS_SYNC();
    if (mode == 1 || mode == 3) r0 = gcs_warp_max(acc_max);
    if (mode == 0 || mode == 4) { r0 = gcs_warp_min(acc_min); r1 = gcs_warp_sum(acc_sum); }
}

// folds the shared gradient slots (y_j, x_i, y_v) written as partials by gcs_rows (modes 0 / 2)
GCS_DEV void gcs_fold(const GcsScratchLayout &L, double *S, int d, bool term, double *gout, int lane) {
    const double *ep = S + L.ep, *cp = S + L.cp, *sy = S + L.sy;
    GCS_LANE_LOOP(j, d) gout[gcs_uw(j) + 4] = ep[GCS_EPP * (2 * j) + 11] + ep[GCS_EPP * (2 * j + 1) + 11] + sy[j];
    GCS_LANE_LOOP(q, 4) {
        int i = q >> 1, c = q & 1;
        double s = cp[8 * i + 2 + c];
        for (int j = 0; j < d; ++j) s += ep[GCS_EPP * (2 * j + i) + 12 + c];
        gout[GCS_UX + q] = s;
    }
    if (lane == 0) {
        gout[GCS_UYV] = cp[1] + cp[8 + 1] + (term ? 0.0 : sy[d]);
        gout[GCS_UT] = 0.0;
    }
    GCS_SYNC();
}

// in-place Cholesky of the n x n matrix H (row stride ldh, lower triangle) with pivot lifting:
// a pivot that falls below the rounding noise of its own cancellation is lifted to that level.
GCS_DEV void gcs_cholesky(double *H, int n, int ldh, double *piv, int lane) {
    for (int j = 0; j < n; ++j) {
        GCS_LANE_LOOP(i, n - j) {
            const int r = j + i;
            const double *hr = H + r * ldh, *hj = H + j * ldh;
            double s = 0.0;
            for (int k = 0; k < j; ++k) s += hr[k] * hj[k];
            if (i == 0) {
                double d0 = hr[j], dd = d0 - s, noise = 64.0 * 2.2e-16 * (fabs(d0) + s) + 1e-300;
                if (!(dd > noise)) dd = noise;
                piv[0] = sqrt(dd);
            } else {
                H[r * ldh + j] = hr[j] - s;
            }
        }
        GCS_SYNC();
        const double sd = piv[0];
        GCS_LANE_LOOP(i, n - j) {
            const int r = j + i;
            if (i == 0) H[r * ldh + j] = sd; else H[r * ldh + j] /= sd;
        }
        GCS_SYNC();
    }
}
// x <- (L L')^-1 x
GCS_DEV void gcs_chol_solve(const double *H, int n, int ldh, double *x, int lane) {
    for (int j = 0; j < n; ++j) {
        const double xj = x[j] / H[j * ldh + j];
        GCS_SYNC();
        if (lane == 0) x[j] = xj;
        GCS_LANE_LOOP(i, n - j - 1) { const int r = j + 1 + i; x[r] -= H[r * ldh + j] * xj; }
        GCS_SYNC();
    }
    for (int j = n - 1; j >= 0; --j) {
        const double xj = x[j] / H[j * ldh + j];
        GCS_SYNC();
        if (lane == 0) x[j] = xj;
        GCS_LANE_LOOP(r, j) x[r] -= H[j * ldh + r] * xj;
        GCS_SYNC();
    }
}

// Solves one vertex program.  Scratch S must already hold: A, b (L.A, L.b), targets (L.tgt, edge-canonical
// order per live half-edge) and the int array out[] (L.ints).  On return S + L.u holds the solution in u-space.
GCS_DEV GcsVertexOut gcs_vertex_solve(const GcsScratchLayout &L, double *S, const GcsVertexIn &in, int lane) {
    const int m = in.m, d = in.d;
    const bool term = in.type != 0;
    const int n = 5 * d, nu = GCS_NCORE + 5 * d, ldh = L.ldh;
    int *out = (int *)(S + L.ints), *prim = out + L.dcap;
    double *A = S + L.A, *b = S + L.b, *AA = S + L.AA, *tgt = S + L.tgt;
    double *u = S + L.u, *dua = S + L.dua, *du = S + L.du, *gu = S + L.gu, *ru = S + L.ru, *Pu = S + L.Pu, *qu = S + L.qu;
    double *v = S + L.v, *dv = S + L.dv, *rv = S + L.rv, *vbest = S + L.vbest;
    double *H = S + L.H, *M = S + L.M, *B = S + L.B, *C = S + L.C;
    GcsVertexOut res; res.iters = 0; res.status = 1; res.gap = 0; res.dres = 0;

    // ---- setup ----------------------------------------------------------------------------------
    int jstar = -1, din = 0, dout = 0;
    for (int j = 0; j < d; ++j) { if (out[j]) dout++; else din++; }
    GCS_LANE_LOOP(j, d) prim[j] = (in.type == 2) ? 1 : out[j];
    GCS_LANE_LOOP(k, m) { AA[3 * k] = A[2 * k] * A[2 * k]; AA[3 * k + 1] = A[2 * k] * A[2 * k + 1]; AA[3 * k + 2] = A[2 * k + 1] * A[2 * k + 1]; }
    for (int j = 0; j < d; ++j) if ((in.type == 2) || out[j]) jstar = j;
    GCS_LANE_LOOP(q, nu) { Pu[q] = 0.0; qu[q] = 0.0; }
    GCS_SYNC();
    const double rho = in.rho;
    GCS_LANE_LOOP(j, d) {   // objective (:380-413): t + eps y + rho/2 |xc - target|^2 on the consensus scalars
        const int o = gcs_uw(j);
        const double *t = tgt + 5 * j;
        const double *T1 = out[j] ? t : t + 2;
        Pu[o] = Pu[o + 1] = rho; qu[o] = -rho * T1[0]; qu[o + 1] = -rho * T1[1];
        if (out[j]) { Pu[o + 2] = Pu[o + 3] = rho; qu[o + 2] = -rho * t[2]; qu[o + 3] = -rho * t[3]; }
        Pu[o + 4] = rho; qu[o + 4] = GCS_EDGE_PENALTY - rho * t[4];
    }
    if (lane == 0) qu[GCS_UT] = 1.0;
    // strictly feasible start (all equalities hold by construction of the map)
    const double eta = term ? 1.0 : 0.5;
    GCS_LANE_LOOP(c, 5) v[c] = (c == 4) ? 1.0 : ((c & 1) ? in.cy : in.cx);
    GCS_LANE_LOOP(q, 5 * (d - 1)) {
        int jj = q / 5, c = q - 5 * jj, j = jj < jstar ? jj : jj + 1;
        double y = eta / (double)(term ? d : (out[j] ? dout : din));
        v[5 + q] = (c == 4) ? y : y * ((c & 1) ? in.cy : in.cx);
    }
    GCS_SYNC();
    gcs_forward(v, u, d, jstar, prim, term, true, lane);
    // slacks -> centred duals  z = mu0 / s  with mu0 = mean slack
    double zq[3], sq[3];
    {
        double part = 0.0; int cnt = 0;
        GCS_LANE_LOOP(e, 2 * d) {
            const int j = e >> 1, i = e & 1, ao = gcs_uw(j) + 2 * i, xo = GCS_UX + 2 * i;
            const double y = u[gcs_uw(j) + 4];
            double *zr = S + L.zr + e * m * 2;
            for (int k = 0; k < m; ++k) {
                const double Aa = A[2 * k] * u[ao] + A[2 * k + 1] * u[ao + 1];
                const double s3 = y * b[k] - Aa;
                zr[2 * k] = s3; part += s3; cnt++;
                if (!term) { const double s4 = (1.0 - y) * b[k] - (A[2 * k] * u[xo] + A[2 * k + 1] * u[xo + 1]) + Aa; zr[2 * k + 1] = s4; part += s4; cnt++; }
                else zr[2 * k + 1] = 1.0;
            }
        }
        GCS_LANE_LOOP(i, 2) {
            const int zo = GCS_UZ + 2 * i, xo = GCS_UX + 2 * i;
            double *zc = S + L.zc + i * m * 2;
            for (int k = 0; k < m; ++k) {
                const double Az = A[2 * k] * u[zo] + A[2 * k + 1] * u[zo + 1];
                const double s1 = u[GCS_UYV] * b[k] - Az;
                zc[2 * k] = s1; part += s1; cnt++;
                if (!term) { const double s2 = (1.0 - u[GCS_UYV]) * b[k] - (A[2 * k] * u[xo] + A[2 * k + 1] * u[xo + 1]) + Az; zc[2 * k + 1] = s2; part += s2; cnt++; }
                else zc[2 * k + 1] = 1.0;
            }
        }
        GCS_LANE_LOOP(j, d + 1) {
            double *zy = S + L.zy;
            if (j < d) { zy[j] = u[gcs_uw(j) + 4]; part += zy[j]; cnt++; }
            else if (!term) { zy[j] = 1.0 - u[GCS_UYV]; part += zy[j]; cnt++; }
            else zy[j] = 1.0;
        }
        GCS_SYNC();
        const double tot = gcs_warp_sum(part), nrows = gcs_warp_sum((double)cnt);
        const double mu0 = tot / nrows;
        GCS_LANE_LOOP(q, 2 * d * m * 2) { double *zr = S + L.zr; zr[q] = mu0 / zr[q]; }
        GCS_LANE_LOOP(q, 2 * m * 2) { double *zc = S + L.zc; zc[q] = mu0 / zc[q]; }
        GCS_LANE_LOOP(j, d + 1) { double *zy = S + L.zy; zy[j] = mu0 / zy[j]; }
        sq[0] = u[GCS_UT]; sq[1] = u[GCS_UZ] - u[GCS_UZ + 2]; sq[2] = u[GCS_UZ + 1] - u[GCS_UZ + 3];
        const double det = gcs_jnorm2(sq[0], sq[1], sq[2]);
        zq[0] = mu0 * sq[0] / det; zq[1] = -mu0 * sq[1] / det; zq[2] = -mu0 * sq[2] / det;
        GCS_SYNC();
    }
    const int nrows_lp = (term ? 1 : 2) * (2 * d * m + 2 * m) + d + (term ? 0 : 1);
    const double deg = (double)(nrows_lp + 1);
    double qn = 1.0;
    {   // scale of the reduced linear cost, for the relative dual residual
        gcs_adjoint(qu, rv, d, jstar, prim, term, lane);
        double part = 0.0;
        GCS_LANE_LOOP(q, n) part = fmax(part, fabs(rv[q]));
        qn = fmax(1.0, gcs_warp_max(part));
        GCS_SYNC();
    }
    GCS_LANE_LOOP(q, n) vbest[q] = v[q];
    double best_merit = 1e300, best_gap = 0.0, best_dres = 0.0;
    int best_it = 0;
    const double tol = in.tol;

    for (int it = 0; it <= in.max_iter; ++it) {
        // ---- residuals and Hessian pieces at the current point --------------------------------
        GCS_LANE_LOOP(q, 100) C[q] = 0.0;
        GCS_SYNC();
        double minprod = 0.0, gap = 0.0;
        { GcsRowsArgs ar; ar.mode = 0; ar.sigmu = 0; ar.alpha = 0; ar.gout = gu; gcs_rows(L, S, m, d, term, ar, minprod, gap, lane); }
        gcs_fold(L, S, d, term, gu, lane);
        sq[0] = u[GCS_UT]; sq[1] = u[GCS_UZ] - u[GCS_UZ + 2]; sq[2] = u[GCS_UZ + 1] - u[GCS_UZ + 3];
        const double pq0 = sq[0] * zq[0] + sq[1] * zq[1] + sq[2] * zq[2];
        gap += pq0; minprod = fmin(minprod, pq0);
        GCS_LANE_LOOP(q, nu) {   // gu = P u + q + G'z - B'zq
            double g = gu[q] + Pu[q] * u[q] + qu[q];
            if (q == GCS_UT) g -= zq[0];
            else if (q == GCS_UZ) g -= zq[1]; else if (q == GCS_UZ + 2) g += zq[1];
            else if (q == GCS_UZ + 1) g -= zq[2]; else if (q == GCS_UZ + 3) g += zq[2];
            gu[q] = g;
        }
        GCS_SYNC();
        gcs_adjoint(gu, rv, d, jstar, prim, term, lane);
        double dres;
        { double part = 0.0; GCS_LANE_LOOP(q, n) part = fmax(part, fabs(rv[q])); dres = gcs_warp_max(part) / qn; }
        res.iters = it; res.gap = gap; res.dres = dres;
#ifdef GCS_EMULATE
        if (getenv("GCSEMU_TRACE")) fprintf(stderr, "   it %2d gap %.3e dres %.3e minprod %.3e sq=(%.2e %.2e %.2e)\n", it, gap, dres, minprod, sq[0], sq[1], sq[2]);
#endif
        if (!(gap == gap) || !(dres == dres)) { res.status = 2; break; }
        if (dres <= 10.0 * tol && gap <= tol) { res.status = 0; break; }
        {
            const double merit = fmax(gap / tol, dres / (10.0 * tol));
            if (merit < best_merit) {
                best_merit = merit; best_gap = gap; best_dres = dres; best_it = it;
                GCS_LANE_LOOP(q, n) vbest[q] = v[q];
            } else if (gap <= tol && it >= best_it + 3) { res.status = 5; break; }   // stalled at the fp64 noise floor
        }
        if (it == in.max_iter) break;
        const double mu = gap / deg;
        GcsNT nt; gcs_nt_build(nt, sq, zq);

        // ---- assemble the blocks of H_u ---------------------------------------------------------
        GCS_LANE_LOOP(j, d) {
            const double *e0 = S + L.ep + GCS_EPP * (2 * j), *e1 = e0 + GCS_EPP;
            double *Mj = M + 25 * j, *Bj = B + 20 * j;
            const int o = gcs_uw(j);
            for (int q = 0; q < 25; ++q) Mj[q] = 0.0;
            for (int q = 0; q < 20; ++q) Bj[q] = 0.0;
            Mj[0] = e0[0] + Pu[o]; Mj[1] = Mj[5] = e0[1]; Mj[6] = e0[2] + Pu[o + 1];
            Mj[12] = e1[0] + Pu[o + 2]; Mj[13] = Mj[17] = e1[1]; Mj[18] = e1[2] + Pu[o + 3];
            Mj[4] = Mj[20] = e0[3]; Mj[9] = Mj[21] = e0[4]; Mj[14] = Mj[22] = e1[3]; Mj[19] = Mj[23] = e1[4];
            Mj[24] = e0[5] + e1[5] + Pu[o + 4] + S[L.dsy + j];
            // B_j: rows x(4), cols w_j(5):  x_i - a_i : Bxa (sym 2x2),  x_i - y : Bxy
            Bj[0] = e0[6]; Bj[1] = e0[7]; Bj[5] = e0[7]; Bj[6] = e0[8]; Bj[4] = e0[9]; Bj[9] = e0[10];
            Bj[12] = e1[6]; Bj[13] = e1[7]; Bj[17] = e1[7]; Bj[18] = e1[8]; Bj[14] = e1[9]; Bj[19] = e1[10];
        }
        GCS_LANE_LOOP(i, 2) {     // x_i x_i gets every C4 row of point i:  sum_j D4 A A' = - sum_j Bxa
            const int X = GCS_UX + 2 * i;
            double s0 = 0, s1 = 0, s2 = 0;
            for (int j = 0; j < d; ++j) { const double *e = S + L.ep + GCS_EPP * (2 * j + i); s0 -= e[6]; s1 -= e[7]; s2 -= e[8]; }
            C[X * 10 + X] += s0; C[X * 10 + X + 1] += s1; C[(X + 1) * 10 + X] += s1; C[(X + 1) * 10 + X + 1] += s2;
        }
        GCS_SYNC();
        if (lane == 0) {
            C[GCS_UYV * 10 + GCS_UYV] = S[L.cp + 0] + S[L.cp + 8] + (term ? 0.0 : S[L.dsy + d]);
            // second-order cone block  B' W^-2 B  on (t, z1 - z2); a few ulps of its trace keep it PSD
            double Wi2[3][3];
            for (int c = 0; c < 3; ++c) {
                double e0 = c == 0, e1 = c == 1, e2 = c == 2, y0, y1, y2, w0, w1, w2;
                gcs_nt_apply(nt, e0, e1, e2, true, y0, y1, y2); gcs_nt_apply(nt, y0, y1, y2, true, w0, w1, w2);
                Wi2[0][c] = w0; Wi2[1][c] = w1; Wi2[2][c] = w2;
            }
            const double lift = 16.0 * 2.2e-16 * (Wi2[0][0] + Wi2[1][1] + Wi2[2][2]);
            Wi2[0][0] += lift; Wi2[1][1] += lift; Wi2[2][2] += lift;
            const int ia[3][2] = {{GCS_UT, -1}, {GCS_UZ, GCS_UZ + 2}, {GCS_UZ + 1, GCS_UZ + 3}};
            for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c)
                for (int sa = 0; sa < 2; ++sa) for (int sc = 0; sc < 2; ++sc) {
                    const int p = ia[a][sa], q = ia[c][sc];
                    if (p < 0 || q < 0) continue;
                    C[p * 10 + q] += ((sa == sc) ? 1.0 : -1.0) * Wi2[a][c];
                }
        }
        GCS_SYNC();
        // ---- H_v = N' H_u N (lower triangle), then factor ------------------------------------
        GCS_LANE_LOOP(idx, n * n) {
            const int p = idx / n, q = idx - p * n;
            if (q > p) continue;
            int ip[3], iq[3]; double cp_[3], cq_[3];
            const int np_ = gcs_ncol(p, jstar, prim, term, ip, cp_), nq_ = gcs_ncol(q, jstar, prim, term, iq, cq_);
            double s = 0.0;
            for (int a = 0; a < np_; ++a) for (int c = 0; c < nq_; ++c) s += cp_[a] * cq_[c] * gcs_hu(M, B, C, ip[a], iq[c]);
            H[p * ldh + q] = s + (p == q ? 1e-14 : 0.0);
        }
        GCS_SYNC();
        gcs_cholesky(H, n, ldh, S + L.cp + 4, lane);

        // ---- predictor:  rhs = -N'(P u + q) -----------------------------------------------------
        GCS_LANE_LOOP(q, nu) ru[q] = -(Pu[q] * u[q] + qu[q]);
        GCS_SYNC();
        gcs_adjoint(ru, dv, d, jstar, prim, term, lane);
        gcs_chol_solve(H, n, ldh, dv, lane);
        gcs_forward(dv, dua, d, jstar, prim, term, false, lane);
        double tmax = 0.0, dummy = 0.0;
        { GcsRowsArgs ar; ar.mode = 1; ar.sigmu = 0; ar.alpha = 0; ar.gout = 0; gcs_rows(L, S, m, d, term, ar, tmax, dummy, lane); }
        double dsq_a[3], dzq_a[3];
        gcs_nt_apply(nt, dua[GCS_UT], dua[GCS_UZ] - dua[GCS_UZ + 2], dua[GCS_UZ + 1] - dua[GCS_UZ + 3], true, dsq_a[0], dsq_a[1], dsq_a[2]);
        dzq_a[0] = -nt.l0 - dsq_a[0]; dzq_a[1] = -nt.l1 - dsq_a[1]; dzq_a[2] = -nt.l2 - dsq_a[2];
        tmax = fmax(tmax, fmax(gcs_soc_max_step(nt, dsq_a), gcs_soc_max_step(nt, dzq_a)));
        const double a_aff = tmax <= 0.0 ? 1.0 : fmin(1.0, 1.0 / tmax);
        double sigma = (1.0 - a_aff) * (1.0 - a_aff) * (1.0 - a_aff);
        {   // centrality-aware floor (LOQO's rule): re-centre when a complementarity product lags
            const double xi = fmax(minprod / mu, 1e-300), c = fmin(0.05 * (1.0 - xi) / xi, 2.0);
            sigma = fmax(sigma, GCS_LOQO_C * c * c * c);
        }
        // ---- corrector --------------------------------------------------------------------------
        { GcsRowsArgs ar; ar.mode = 2; ar.sigmu = sigma * mu; ar.alpha = 0; ar.gout = ru; gcs_rows(L, S, m, d, term, ar, dummy, dummy, lane); }
        gcs_fold(L, S, d, term, ru, lane);
        double tq[3];
        {
            double dsoc[3];
            dsoc[0] = -(nt.l0 * nt.l0 + nt.l1 * nt.l1 + nt.l2 * nt.l2) + sigma * mu - (dsq_a[0] * dzq_a[0] + dsq_a[1] * dzq_a[1] + dsq_a[2] * dzq_a[2]);
            dsoc[1] = -2.0 * nt.l0 * nt.l1 - (dsq_a[0] * dzq_a[1] + dzq_a[0] * dsq_a[1]);
            dsoc[2] = -2.0 * nt.l0 * nt.l2 - (dsq_a[0] * dzq_a[2] + dzq_a[0] * dsq_a[2]);
            gcs_soc_div(nt, dsoc, tq);
        }
        {
            double w0, w1, w2; gcs_nt_apply(nt, tq[0], tq[1], tq[2], true, w0, w1, w2);
            const double sc = 1.0 - sigma;
            GCS_LANE_LOOP(q, nu) {
                double r = ru[q] - sc * gu[q];
                if (q == GCS_UT) r += w0;
                else if (q == GCS_UZ) r += w1; else if (q == GCS_UZ + 2) r -= w1;
                else if (q == GCS_UZ + 1) r += w2; else if (q == GCS_UZ + 3) r -= w2;
                ru[q] = r;
            }
            GCS_SYNC();
        }
        gcs_adjoint(ru, dv, d, jstar, prim, term, lane);
        gcs_chol_solve(H, n, ldh, dv, lane);
        gcs_forward(dv, du, d, jstar, prim, term, false, lane);
        { GcsRowsArgs ar; ar.mode = 3; ar.sigmu = sigma * mu; ar.alpha = 0; ar.gout = 0; gcs_rows(L, S, m, d, term, ar, tmax, dummy, lane); }
        double dsq_s[3], dzq_s[3];
        gcs_nt_apply(nt, du[GCS_UT], du[GCS_UZ] - du[GCS_UZ + 2], du[GCS_UZ + 1] - du[GCS_UZ + 3], true, dsq_s[0], dsq_s[1], dsq_s[2]);
        dzq_s[0] = tq[0] - dsq_s[0]; dzq_s[1] = tq[1] - dsq_s[1]; dzq_s[2] = tq[2] - dsq_s[2];
        tmax = fmax(tmax, fmax(gcs_soc_max_step(nt, dsq_s), gcs_soc_max_step(nt, dzq_s)));
        double alpha = tmax <= 0.0 ? 1.0 : fmin(1.0, 0.99 / tmax);
        double dzq[3], dsq[3];
        gcs_nt_apply(nt, dzq_s[0], dzq_s[1], dzq_s[2], true, dzq[0], dzq[1], dzq[2]);
        gcs_nt_apply(nt, dsq_s[0], dsq_s[1], dsq_s[2], false, dsq[0], dsq[1], dsq[2]);
        for (int bt = 0; bt < 30; ++bt) {    // wide-neighbourhood safeguard
            double mn = 0.0, sum = 0.0;
            { GcsRowsArgs ar; ar.mode = 4; ar.sigmu = 0; ar.alpha = alpha; ar.gout = 0; gcs_rows(L, S, m, d, term, ar, mn, sum, lane); }
            double pq = 0.0;
            for (int k = 0; k < 3; ++k) pq += (sq[k] + alpha * dsq[k]) * (zq[k] + alpha * dzq[k]);
            sum += pq; mn = fmin(mn, pq);
            if (mn >= GCS_NB_GAMMA * sum / deg) break;
            alpha *= 0.7;
        }
        {   // a non-finite direction ends the solve on the best iterate seen
            double bad = 0.0;
            GCS_LANE_LOOP(q, n) { const double x = dv[q]; if (!(x == x) || fabs(x) > 1e300) bad = 1.0; }
            bad = gcs_warp_max(bad);
            if (bad > 0.0 || !(alpha == alpha)) { res.status = 2; break; }
        }
        GCS_LANE_LOOP(q, n) v[q] += alpha * dv[q];
        { GcsRowsArgs ar; ar.mode = 5; ar.sigmu = 0; ar.alpha = alpha; ar.gout = 0; gcs_rows(L, S, m, d, term, ar, dummy, dummy, lane); }
        zq[0] += alpha * dzq[0]; zq[1] += alpha * dzq[1]; zq[2] += alpha * dzq[2];
        GCS_SYNC();
        gcs_forward(v, u, d, jstar, prim, term, true, lane);
    }
    if (res.status != 0) {
        GCS_LANE_LOOP(q, n) v[q] = vbest[q];
        GCS_SYNC();
        gcs_forward(v, u, d, jstar, prim, term, true, lane);
        if (best_merit < 1e300) { res.gap = best_gap; res.dres = best_dres; }
        // the best iterate sits at the fp64 noise floor of this program: accept it
        if (best_gap <= 10.0 * tol && best_dres <= 100.0 * tol) res.status = 5;
    }
    return res;
}
