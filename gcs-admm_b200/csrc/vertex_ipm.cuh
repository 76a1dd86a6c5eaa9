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
#define GCS_DEV_NI static inline
#define GCS_LANE_LOOP(i, n) for (int i = 0; i < (n); ++i)
#define GCS_SYNC() ((void)0)
static inline double gcs_warp_sum(double x) { return x; }
static inline double gcs_warp_max(double x) { return x; }
static inline double gcs_warp_min(double x) { return x; }
#else
#define GCS_DEV __device__ __forceinline__
#define GCS_DEV_NI __device__ __noinline__   // called several times per iteration: one copy keeps the kernel in the I-cache
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
#define GCS_EHP 6              // Hessian part of a row-family record: sum D AA' (3), sum D b A (2), sum D b^2
#define GCS_EGP 3              // gradient part: sum w A (2), sum w b
#define GCS_NB_GAMMA 1e-5      // width of the central-path neighbourhood
#define GCS_LOQO_C 0.02        // weight of the centrality-aware floor on sigma
#define GCS_QN 21               // doubles per compact Hessian block: aa0(3) aa1(3) ay0(2) ay1(2) yy xa0(3) xa1(3) xy0(2) xy1(2)

// Scratch layout (offsets in doubles) of one warp, for <= dcap live half-edges and <= mcap polytope rows.
struct GcsScratchLayout {
    int dcap, mcap, ncap, nucap;
    int H, C, u, gu, Pu, qu, dua, du, ru, dv, rv, ytmp, Q, v, vbest, zr, wr, zy, dsy, dzy, sy;
    int eh, eg, A, b, AA, tgt, ints, diag0, Linv, MS, CZ, total;
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
    int o = 0;
    // Arrays that are never live together share storage (21.7 -> 18.8 KB per warp for d = 8, m = 8: 12 instead of 10 warps per SM):
    //   eh (Hessian records of the row families, written by the first row pass) | H (assembled from Q, C after eh has been folded into them)
    //   C (core block, assembly only) | Linv (written by the factorisation);  MS | ytmp (solve temporary);  CZ | rv (residual check only)
    {
        const int hsz = L.ncap * (L.ncap + 1) / 2, esz = GCS_EHP * 4 * (dcap + 1);
        L.H = o; L.eh = o; o += hsz > esz ? hsz : esz;      // H: packed lower triangle, row r starts at r (r + 1) / 2
    }
    L.u = o; o += L.nucap;  L.gu = o; o += L.nucap;  L.Pu = o; o += L.nucap;  L.qu = o; o += L.nucap;
    // direction / right-hand-side vectors; between the residual check and the factorisation they are
    // free and hold Q, the compact Hessian blocks (21 doubles per half-edge block)
    L.dua = o; L.Q = o; o += L.nucap;  L.du = o; o += L.nucap;  L.ru = o; o += L.nucap;
    L.dv = o; o += L.ncap;
    L.rv = o; L.CZ = o; o += L.ncap > 25 ? L.ncap : 25;
    L.ytmp = o; L.MS = o; o += L.ncap > 25 ? L.ncap : 25;
    L.v = o; o += L.ncap;    L.vbest = o; o += L.ncap;
    int nr = 4 * (dcap + 1) * mcap;            // one slot of mcap rows per family (block, point, kind); block dcap = core
    L.zr = o; o += nr;  L.wr = o; o += nr;      // duals; work array (1/s, then dz)
    L.zy = o; o += dcap + 1; L.dsy = o; o += dcap + 1; L.dzy = o; o += dcap + 1; L.sy = o; o += dcap + 1;
    L.eg = o; o += GCS_EGP * 4 * (dcap + 1);     // gradient records (rewritten by the corrector pass while H holds the factor)
    L.A = o; o += 2 * mcap; L.b = o; o += mcap; L.AA = o; o += 3 * mcap;
    L.tgt = o; o += 5 * dcap;
    L.ints = o; o += (3 * dcap + 1) / 2 + 1;   // int arrays out / prim / hid packed behind the doubles
    L.diag0 = o; o += L.ncap;
    L.Linv = o; L.C = o; o += 15 * dcap > 100 ? 15 * dcap : 100;
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
    double *ws;            // warm-start record of this vertex in global memory (gcs_ws_stride doubles) or null
    double theta;          // warm-start mixing weight towards the analytic centre (0 = cold start)
};

// warm-start record: [valid][v (ncap)][row duals (4 (dcap+1) mcap)][single duals (dcap+1)][cone dual (3)]
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int gcs_ws_stride(const GcsScratchLayout &L) { return 1 + L.ncap + 4 * (L.dcap + 1) * L.mcap + (L.dcap + 1) + 3; }

struct GcsVertexOut { int iters; int status; double gap, dres; };

struct GcsNT { double w0, w1, w2, beta, l0, l1, l2, lb0, lb1, lb2, inrm; };   // lb = lam / |lam|_J, inrm = 1 / |lam|_J

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
    S.inrm = 1.0 / sqrt(fmax(gcs_jnorm2(S.l0, S.l1, S.l2), 1e-300));
    S.lb0 = S.l0 * S.inrm; S.lb1 = S.l1 * S.inrm; S.lb2 = S.l2 * S.inrm;
}
GCS_DEV double gcs_soc_max_step(const GcsNT &S, const double *d) {
    const double c0 = S.lb0 * d[0] - S.lb1 * d[1] - S.lb2 * d[2], f = (c0 + d[0]) / (S.lb0 + 1.0);
    return (hypot(d[1] - f * S.lb1, d[2] - f * S.lb2) - c0) * S.inrm;
}
GCS_DEV void gcs_soc_div(const GcsNT &S, const double *v, double *x) {  // lam o x = v
    double det = gcs_jnorm2(S.l0, S.l1, S.l2);
    double x0 = (S.l0 * v[0] - S.l1 * v[1] - S.l2 * v[2]) / det;
    x[0] = x0; x[1] = (v[1] - x0 * S.l1) / S.l0; x[2] = (v[2] - x0 * S.l2) / S.l0;
}

GCS_DEV int gcs_uw(int j) { return GCS_NCORE + 5 * j; }   // u-space offset of live half-edge j

// ---- the null-space map ---------------------------------------------------------------------
// forward:  u = N v (+ up when `affine`)
GCS_DEV_NI void gcs_forward(const double *v, double *u, int d, int jstar, const int *prim, bool term, bool affine, int lane) {
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
GCS_DEV_NI void gcs_adjoint(const double *g, double *out, int d, int jstar, const int *prim, bool term, int lane) {
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

#ifdef GCS_EMULATE
static inline double gcs_rcp(double x) { return 1.0 / x; }
#else
__device__ __forceinline__ double gcs_rcp(double x) { return __drcp_rn(x); }
#endif

// ---- one pass over the inequality rows ---------------------------------------------------------
// Every polytope row family has the form  s_k = h0 b_k - A_k . p  (k = 0..m-1):
//   C3 (:434-436)  h0 = y_e^v      p = a_i             C1 (:420-422)  h0 = y_v      p = z_i
//   C4 (:438-440)  h0 = 1 - y_e^v  p = x_i - a_i       C2 (:424-426)  h0 = 1 - y_v  p = x_i - z_i
// One lane owns one family (block, point i, kind): 4 (d + 1) items (2 (d + 1) for 's'/'t', which have no
// C2 / C4).  Each item leaves a record  [sum D AA'(3), sum D b A(2), sum D b^2, sum w A(2), sum w b]
// (D = z/s; w = z in mode 0, rc/s in mode 2) that the assembly signs into M_j, B_j, C and the gradient.
//   mode 0: records, gap (r1) and smallest complementarity product (r0); caches 1/s
//   mode 1: predictor step ratio for direction dua (r0)
//   mode 2: corrector right-hand side records (rc uses dua and sigmu)
//   mode 3: corrector direction du: store ds, dz per row; r0 = max ratio
//   mode 4: neighbourhood statistics for step alpha: r0 = min, r1 = sum of (s + a ds)(z + a dz)
//   mode 5: z += alpha dz
//   mode 6: initial slacks into the dual slots; r1 = their sum, r0 = their count
struct GcsRowsArgs { int mode; double sigmu, alpha; };

GCS_DEV int gcs_slot(int blk, int i, int fam) { return (blk * 2 + i) * 2 + fam; }

GCS_DEV void gcs_rows(const GcsScratchLayout &L, double *S, int m, int d, bool term, const GcsRowsArgs &ar,
                      double &r0, double &r1, int lane) {
    const double *__restrict__ A = S + L.A, *__restrict__ b = S + L.b, *__restrict__ AA = S + L.AA;
    const double *__restrict__ u = S + L.u, *__restrict__ du = S + L.du, *__restrict__ dua = S + L.dua;
    double acc_max = 0.0, acc_min = 1e300, acc_sum = 0.0, acc_cnt = 0.0;
    const int mode = ar.mode;
    const bool need_p = (mode == 1 || mode == 2 || mode == 3), need_d = (mode == 3 || mode == 4);
    const int nitems = term ? 2 * (d + 1) : 4 * (d + 1);
    GCS_LANE_LOOP(it, nitems) {
        int fam, i, blk;
        if (term) { fam = 0; i = it & 1; blk = it >> 1; } else { fam = it & 1; i = (it >> 1) & 1; blk = it >> 2; }
        const int po = (blk < d ? gcs_uw(blk) : GCS_UZ) + 2 * i, yo = blk < d ? gcs_uw(blk) + 4 : GCS_UYV, xo = GCS_UX + 2 * i;
        const int slot = gcs_slot(blk, i, fam);
        const double h0 = fam ? 1.0 - u[yo] : u[yo];
        const double p0 = fam ? u[xo] - u[po] : u[po], p1 = fam ? u[xo + 1] - u[po + 1] : u[po + 1];
        double ph = 0, pp0 = 0, pp1 = 0, dh = 0, dp0 = 0, dp1 = 0;
        if (need_p) { ph = fam ? -dua[yo] : dua[yo]; pp0 = fam ? dua[xo] - dua[po] : dua[po]; pp1 = fam ? dua[xo + 1] - dua[po + 1] : dua[po + 1]; }
        if (need_d) { dh = fam ? -du[yo] : du[yo]; dp0 = fam ? du[xo] - du[po] : du[po]; dp1 = fam ? du[xo + 1] - du[po + 1] : du[po + 1]; }
        // row-major over k, slot-minor: lanes (= slots) touch consecutive doubles, no shared-memory bank conflicts
        const int NS = 4 * (L.dcap + 1);
        double *__restrict__ zr = S + L.zr + slot, *__restrict__ wr = S + L.wr + slot;
        double s_aa0 = 0, s_aa1 = 0, s_aa2 = 0, s_ba0 = 0, s_ba1 = 0, s_bb = 0, w_a0 = 0, w_a1 = 0, w_b = 0;
        for (int k = 0; k < m; ++k) {
            const double A0 = A[2 * k], A1 = A[2 * k + 1], bk = b[k];
            const double sl = h0 * bk - (A0 * p0 + A1 * p1);
            if (mode == 6) { zr[k * NS] = sl; acc_sum += sl; acc_cnt += 1.0; continue; }
            const double z = zr[k * NS];
            if (mode == 0) {
                const double rs = gcs_rcp(sl), D = z * rs;
                wr[k * NS] = rs;                                   // cached 1/s for the other passes of this iteration
                s_aa0 += D * AA[3 * k]; s_aa1 += D * AA[3 * k + 1]; s_aa2 += D * AA[3 * k + 2];
                s_ba0 += D * bk * A0; s_ba1 += D * bk * A1; s_bb += D * bk * bk;
                w_a0 += A0 * z; w_a1 += A1 * z; w_b += bk * z;
                const double pr = sl * z;
                acc_sum += pr; acc_min = fmin(acc_min, pr);
            } else if (need_p) {
                const double rs = wr[k * NS];
                const double ps = ph * bk - (A0 * pp0 + A1 * pp1);      // predictor ds
                const double t = ps * rs;                                // ds / s
                if (mode == 1) {
                    acc_max = fmax(acc_max, fmax(-t, 1.0 + t));          // -ds/s and -dz/z = 1 + ds/s  (dz = -z - z ds/s)
                } else {
                    const double pz = -z - z * t;
                    const double rc = -sl * z + ar.sigmu - ps * pz;
                    if (mode == 2) {
                        const double g = rc * rs;
                        w_a0 += A0 * g; w_a1 += A1 * g; w_b += bk * g;
                    } else {
                        const double ds = dh * bk - (A0 * dp0 + A1 * dp1);
                        const double dz = (rc - z * ds) * rs;
                        wr[k * NS] = dz;                           // 1/s is not needed after this pass
                        acc_max = fmax(acc_max, fmax(-ds * rs, -dz * gcs_rcp(z)));
                    }
                }
            } else if (mode == 4) {
                const double ds = dh * bk - (A0 * dp0 + A1 * dp1);
                const double pr = (sl + ar.alpha * ds) * (z + ar.alpha * wr[k * NS]);
                acc_sum += pr; acc_min = fmin(acc_min, pr);
            } else {
                zr[k * NS] = z + ar.alpha * wr[k * NS];
            }
        }
        if (mode == 0 || mode == 2) {
            if (mode == 0) { double *rec = S + L.eh + GCS_EHP * slot; rec[0] = s_aa0; rec[1] = s_aa1; rec[2] = s_aa2; rec[3] = s_ba0; rec[4] = s_ba1; rec[5] = s_bb; }
            double *rg = S + L.eg + GCS_EGP * slot;
            rg[0] = w_a0; rg[1] = w_a1; rg[2] = w_b;
        }
    }
    // ---- singles: y_e >= 0 (:377) and y_v <= 1 (:366)
    GCS_LANE_LOOP(j, d + 1) {
        double *zy = S + L.zy, *dsy = S + L.dsy, *dzy = S + L.dzy, *sy = S + L.sy;
        if (j == d && term) { if (mode == 0 || mode == 2) sy[j] = 0.0; if (mode == 0) dsy[j] = 0.0; continue; }
        const int yo = j < d ? gcs_uw(j) + 4 : GCS_UYV;
        const double sgn = j < d ? 1.0 : -1.0;           // s = y   or   s = 1 - y_v ;  row g = -sgn on y
        const double sl = j < d ? u[yo] : 1.0 - u[yo];
        if (mode == 6) { zy[j] = sl; acc_sum += sl; acc_cnt += 1.0; continue; }
        const double z = zy[j];
        if (mode == 0) {
            const double pr = sl * z;
            acc_sum += pr; acc_min = fmin(acc_min, pr);
            dsy[j] = z / sl;         // D, consumed by the assembly
            sy[j] = -sgn * z;        // G'z contribution
        } else if (need_p) {
            const double ps = sgn * dua[yo], pz = -z - z * ps / sl;
            if (mode == 1) acc_max = fmax(acc_max, fmax(-ps / sl, -pz / z));
            else {
                const double rc = -sl * z + ar.sigmu - ps * pz;
                if (mode == 2) sy[j] = sgn * rc / sl;     // -G'(rc/s)
                else {
                    const double ds = sgn * du[yo], dz = (rc - z * ds) / sl;
                    dsy[j] = ds; dzy[j] = dz;
                    acc_max = fmax(acc_max, fmax(-ds / sl, -dz / z));
                }
            }
        } else if (mode == 4) {
            const double pr = (sl + ar.alpha * dsy[j]) * (z + ar.alpha * dzy[j]);
            acc_sum += pr; acc_min = fmin(acc_min, pr);
        } else {
            zy[j] = z + ar.alpha * dzy[j];
        }
    }
    GCS_SYNC();
    if (mode == 1 || mode == 3) r0 = gcs_warp_max(acc_max);
    if (mode == 0 || mode == 4) { r0 = gcs_warp_min(acc_min); r1 = gcs_warp_sum(acc_sum); }
    if (mode == 6) { r0 = gcs_warp_sum(acc_cnt); r1 = gcs_warp_sum(acc_sum); }
}

// gradient-like vector from the records:  gout = sign * G'w (+ the singles' signed terms)
//   a_i / z_i slots:  wA(C3|C1) - wA(C4|C2);   y / y_v slots:  -wb(C3|C1) + wb(C4|C2);   x_i: sum over blocks of wA(C4|C2)
GCS_DEV_NI void gcs_fold(const GcsScratchLayout &L, double *S, int d, bool term, double *gout, double sign, int lane) {
    const double *eg = S + L.eg, *sy = S + L.sy;
    GCS_LANE_LOOP(q, 4 * (d + 1)) {
        const int blk = q >> 2, i = (q >> 1) & 1, c = q & 1;
        const double *r3 = eg + GCS_EGP * gcs_slot(blk, i, 0);
        const double v = r3[c] - (term ? 0.0 : r3[GCS_EGP + c]);
        gout[(blk < d ? gcs_uw(blk) : GCS_UZ) + 2 * i + c] = sign * v;
    }
    GCS_LANE_LOOP(blk, d + 1) {
        const double *r = eg + GCS_EGP * gcs_slot(blk, 0, 0);
        double v = -(r[2] + r[2 * GCS_EGP + 2]);
        if (!term) v += r[GCS_EGP + 2] + r[3 * GCS_EGP + 2];
        gout[blk < d ? gcs_uw(blk) + 4 : GCS_UYV] = sign * v + sy[blk];
    }
    GCS_LANE_LOOP(q, 4) {
        const int i = q >> 1, c = q & 1;
        double v = 0.0;
        if (!term) for (int blk = 0; blk <= d; ++blk) v += eg[GCS_EGP * gcs_slot(blk, i, 1) + c];
        gout[GCS_UX + q] = sign * v;
    }
    if (lane == 0) gout[GCS_UT] = 0.0;
    GCS_SYNC();
}

#ifdef GCS_EMULATE
static inline double gcs_rsqrt(double x) { return 1.0 / sqrt(x); }
#else
__device__ __forceinline__ double gcs_rsqrt(double x) { return rsqrt(x); }
#endif

// In-place blocked Cholesky (block size 5 = one half-edge block) of the n x n matrix H, n = 5 nb,
// row stride ldh, lower triangle.  Pivot lifting: a pivot that falls below the rounding noise of its
// own cancellation is lifted to that level (diag0 holds the diagonal before elimination).
// Linv[b] receives the inverse of the b-th diagonal Cholesky block (15 doubles, row-major lower).
GCS_DEV int gcs_tri(int r) { return (r * (r + 1)) >> 1; }   // start of row r in the packed lower triangle

GCS_DEV void gcs_cholesky(double *H, int nb, const double *diag0, double *Linv, int lane) {
    const int n = 5 * nb;
    for (int b = 0; b < nb; ++b) {
        const int o = 5 * b;
        // (1) every lane factors the 5x5 diagonal block redundantly in registers
        double l[15], li[15];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) l[r * (r + 1) / 2 + c] = H[gcs_tri(o + r) + o + c];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const double d0 = diag0[o + j];
            double dd = l[j * (j + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) dd -= l[j * (j + 1) / 2 + k] * l[j * (j + 1) / 2 + k];
            const double noise = 64.0 * 2.2e-16 * (fabs(d0) + fabs(d0 - dd)) + 1e-300;
            if (!(dd > noise)) dd = noise;
            const double rs = gcs_rsqrt(dd);
            l[j * (j + 1) / 2 + j] = dd * rs;
            li[j * (j + 1) / 2 + j] = rs;
#pragma unroll
            for (int r = j + 1; r < 5; ++r) {
                double v = l[r * (r + 1) / 2 + j];
#pragma unroll
                for (int k = 0; k < j; ++k) v -= l[r * (r + 1) / 2 + k] * l[j * (j + 1) / 2 + k];
                l[r * (r + 1) / 2 + j] = v * rs;
            }
        }
        // inverse of the lower-triangular block
#pragma unroll
        for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int r = c + 1; r < 5; ++r) {
                double v = 0.0;
#pragma unroll
                for (int k = c; k < r; ++k) v -= l[r * (r + 1) / 2 + k] * li[k * (k + 1) / 2 + c];
                li[r * (r + 1) / 2 + c] = v * li[r * (r + 1) / 2 + r];
            }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) { H[gcs_tri(o + r) + o + c] = l[r * (r + 1) / 2 + c]; Linv[15 * b + r * (r + 1) / 2 + c] = li[r * (r + 1) / 2 + c]; }
        }
        // (2) panel: rows below the block,  L_panel = H_panel L_bb^-T
        const int rem = n - o - 5;
        GCS_LANE_LOOP(i, rem) {
            double *hr = H + gcs_tri(o + 5 + i) + o;
            const double x0 = hr[0], x1 = hr[1], x2 = hr[2], x3 = hr[3], x4 = hr[4];
            hr[0] = x0 * li[0];
            hr[1] = x0 * li[1] + x1 * li[2];
            hr[2] = x0 * li[3] + x1 * li[4] + x2 * li[5];
            hr[3] = x0 * li[6] + x1 * li[7] + x2 * li[8] + x3 * li[9];
            hr[4] = x0 * li[10] + x1 * li[11] + x2 * li[12] + x3 * li[13] + x4 * li[14];
        }
        GCS_SYNC();
        // (3) trailing update; long rows first so the last partial round holds the short ones.
        //     Row r only writes its own entries and only reads the (already final) panel columns, so the loads
        //     of four steps are issued before the four stores.
        GCS_LANE_LOOP(i, rem) {
            const int r = n - 1 - i;
            const double *pr = H + gcs_tri(r) + o;
            const double p0 = pr[0], p1 = pr[1], p2 = pr[2], p3 = pr[3], p4 = pr[4];
            double *hrow = H + gcs_tri(r);
            int k = o + 5;
            for (; k + 3 <= r; k += 4) {
                const double *q0 = H + gcs_tri(k) + o, *q1 = H + gcs_tri(k + 1) + o, *q2 = H + gcs_tri(k + 2) + o, *q3 = H + gcs_tri(k + 3) + o;
                const double s0 = p0 * q0[0] + p1 * q0[1] + p2 * q0[2] + p3 * q0[3] + p4 * q0[4];
                const double s1 = p0 * q1[0] + p1 * q1[1] + p2 * q1[2] + p3 * q1[3] + p4 * q1[4];
                const double s2 = p0 * q2[0] + p1 * q2[1] + p2 * q2[2] + p3 * q2[3] + p4 * q2[4];
                const double s3 = p0 * q3[0] + p1 * q3[1] + p2 * q3[2] + p3 * q3[3] + p4 * q3[4];
                const double h0 = hrow[k], h1 = hrow[k + 1], h2 = hrow[k + 2], h3 = hrow[k + 3];
                hrow[k] = h0 - s0; hrow[k + 1] = h1 - s1; hrow[k + 2] = h2 - s2; hrow[k + 3] = h3 - s3;
            }
            for (; k <= r; ++k) {
                const double *pk = H + gcs_tri(k) + o;
                hrow[k] -= p0 * pk[0] + p1 * pk[1] + p2 * pk[2] + p3 * pk[3] + p4 * pk[4];
            }
        }
        GCS_SYNC();
    }
}
// x <- (L L')^-1 x   (y: scratch of n doubles)
GCS_DEV_NI void gcs_chol_solve(const double *__restrict__ H, int nb, const double *__restrict__ Linv, double *__restrict__ x, double *__restrict__ y, int lane) {
    const int n = 5 * nb;
    for (int b = 0; b < nb; ++b) {          // forward: L y = x
        const int o = 5 * b;
        const double *li = Linv + 15 * b;
        const double x0 = x[o], x1 = x[o + 1], x2 = x[o + 2], x3 = x[o + 3], x4 = x[o + 4];
        const double y0 = x0 * li[0], y1 = x0 * li[1] + x1 * li[2], y2 = x0 * li[3] + x1 * li[4] + x2 * li[5];
        const double y3 = x0 * li[6] + x1 * li[7] + x2 * li[8] + x3 * li[9];
        const double y4 = x0 * li[10] + x1 * li[11] + x2 * li[12] + x3 * li[13] + x4 * li[14];
        if (lane == 0) { y[o] = y0; y[o + 1] = y1; y[o + 2] = y2; y[o + 3] = y3; y[o + 4] = y4; }
        GCS_LANE_LOOP(i, n - o - 5) {
            const int r = o + 5 + i;
            const double *hr = H + gcs_tri(r) + o;
            x[r] -= hr[0] * y0 + hr[1] * y1 + hr[2] * y2 + hr[3] * y3 + hr[4] * y4;
        }
        GCS_SYNC();
    }
    for (int b = nb - 1; b >= 0; --b) {     // backward: L' x = y
        const int o = 5 * b;
        const double *li = Linv + 15 * b;
        const double y0 = y[o], y1 = y[o + 1], y2 = y[o + 2], y3 = y[o + 3], y4 = y[o + 4];
        const double x4 = y4 * li[14];
        const double x3 = y3 * li[9] + y4 * li[13];
        const double x2 = y2 * li[5] + y3 * li[8] + y4 * li[12];
        const double x1 = y1 * li[2] + y2 * li[4] + y3 * li[7] + y4 * li[11];
        const double x0 = y0 * li[0] + y1 * li[1] + y2 * li[3] + y3 * li[6] + y4 * li[10];
        if (lane == 0) { x[o] = x0; x[o + 1] = x1; x[o + 2] = x2; x[o + 3] = x3; x[o + 4] = x4; }
        GCS_LANE_LOOP(q, o) {
            y[q] -= H[gcs_tri(o) + q] * x0 + H[gcs_tri(o + 1) + q] * x1 + H[gcs_tri(o + 2) + q] * x2 + H[gcs_tri(o + 3) + q] * x3 + H[gcs_tri(o + 4) + q] * x4;
        }
        GCS_SYNC();
    }
}

// compact Hessian block Q[21] = aa0(3) aa1(3) ay0(2) ay1(2) yy xa0(3) xa1(3) xy0(2) xy1(2)
//   M(a, c): the block's own 5x5 (a1 a2 y);   B(xr, c): its coupling to x (rows x1 x2)
GCS_DEV double gcs_qM(const double *Qj, int a, int c) {
    if (a == 4 && c == 4) return Qj[10];
    if (a == 4 || c == 4) return Qj[6 + (a < c ? a : c)];
    if ((a >> 1) != (c >> 1)) return 0.0;
    return Qj[3 * (a >> 1) + (a & 1) + (c & 1)];
}
GCS_DEV double gcs_qB(const double *Qj, int xr, int c) {
    if (c == 4) return Qj[17 + xr];
    if ((xr >> 1) != (c >> 1)) return 0.0;
    return Qj[11 + 3 * (xr >> 1) + (xr & 1) + (c & 1)];
}

// Solves one vertex program.  Scratch S must already hold: A, b (L.A, L.b), targets (L.tgt, edge-canonical
// order per live half-edge) and the int array out[] (L.ints).  On return S + L.u holds the solution in u-space.
GCS_DEV GcsVertexOut gcs_vertex_solve(const GcsScratchLayout &L, double *S, const GcsVertexIn &in, int lane) {
    const int m = in.m, d = in.d;
    const bool term = in.type != 0;
    const int n = 5 * d, nu = GCS_NCORE + 5 * d;
    int *out = (int *)(S + L.ints), *prim = out + L.dcap;
    double *A = S + L.A, *b = S + L.b, *AA = S + L.AA, *tgt = S + L.tgt;
    double *u = S + L.u, *dua = S + L.dua, *du = S + L.du, *gu = S + L.gu, *ru = S + L.ru, *Pu = S + L.Pu, *qu = S + L.qu;
    double *v = S + L.v, *dv = S + L.dv, *rv = S + L.rv, *vbest = S + L.vbest;
    double *H = S + L.H, *Q = S + L.Q, *C = S + L.C;
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
    // slacks -> centred duals  z = mu0 / s  with mu0 = mean slack.
    // Warm start (the feasible set does not change between ADMM iterations, only the targets do): pull the
    // previous optimum a fraction theta towards the analytic centre — strictly feasible by convexity — keep
    // its duals and add the centring term  theta * mean-slack / s  so every complementarity product is positive.
    double zq[3], sq[3];
    {
        const int nrw = 4 * (L.dcap + 1) * L.mcap;
        const bool warm = in.ws != 0 && in.theta > 0.0 && in.ws[0] == 1.0;
        const double *wv = in.ws + 1, *wz = in.ws + 1 + L.ncap, *wy = wz + nrw, *wq = wy + L.dcap + 1;
        if (warm) {
            GCS_LANE_LOOP(q, n) v[q] = wv[q] + in.theta * (v[q] - wv[q]);
            GCS_SYNC();
            gcs_forward(v, u, d, jstar, prim, term, true, lane);
        }
        double cnt = 0.0, tot = 0.0;
        { GcsRowsArgs ar; ar.mode = 6; ar.sigmu = 0; ar.alpha = 0; gcs_rows(L, S, m, d, term, ar, cnt, tot, lane); }
        double mu0 = tot / cnt;
#ifdef GCS_EMULATE
        if (getenv("GCSEMU_MU0")) mu0 *= atof(getenv("GCSEMU_MU0"));
#endif
        if (warm) mu0 *= in.theta;
        {
            const int NS = 4 * (L.dcap + 1), ns = 4 * (d + 1);
            GCS_LANE_LOOP(q, ns * m) {         // (k, slot) -> k * NS + slot, in scratch and in the record alike
                const int k = q / ns, idx = k * NS + (q - k * ns);
                S[L.zr + idx] = mu0 / S[L.zr + idx] + (warm ? wz[idx] : 0.0);
            }
        }
        GCS_LANE_LOOP(j, d + 1) { double *zy = S + L.zy; if (!(j == d && term)) zy[j] = mu0 / zy[j] + (warm ? wy[j] : 0.0); }
        sq[0] = u[GCS_UT]; sq[1] = u[GCS_UZ] - u[GCS_UZ + 2]; sq[2] = u[GCS_UZ + 1] - u[GCS_UZ + 3];
        const double det = gcs_jnorm2(sq[0], sq[1], sq[2]);
        zq[0] = mu0 * sq[0] / det; zq[1] = -mu0 * sq[1] / det; zq[2] = -mu0 * sq[2] / det;
        if (warm) { zq[0] += wq[0]; zq[1] += wq[1]; zq[2] += wq[2]; }
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
        { GcsRowsArgs ar; ar.mode = 0; ar.sigmu = 0; ar.alpha = 0; gcs_rows(L, S, m, d, term, ar, minprod, gap, lane); }
        gcs_fold(L, S, d, term, gu, 1.0, lane);
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

        // ---- assemble the blocks of H_u from the family records ---------------------------------
        // block (a_i | z_i, y | y_v) gets  +S_AA, -S_bA, +S_bb  from both kinds; the x_i coupling comes from C4 | C2 only
        GCS_LANE_LOOP(blk, d + 1) {
            const double *r03 = S + L.eh + GCS_EHP * gcs_slot(blk, 0, 0), *r04 = r03 + GCS_EHP, *r13 = r03 + 2 * GCS_EHP, *r14 = r03 + 3 * GCS_EHP;
            double aa0[3], aa1[3], ay0[2], ay1[2], yy, xa0[3], xa1[3], xy0[2], xy1[2];
            for (int q = 0; q < 3; ++q) { aa0[q] = r03[q] + (term ? 0.0 : r04[q]); aa1[q] = r13[q] + (term ? 0.0 : r14[q]); xa0[q] = term ? 0.0 : -r04[q]; xa1[q] = term ? 0.0 : -r14[q]; }
            for (int q = 0; q < 2; ++q) { ay0[q] = -(r03[3 + q] + (term ? 0.0 : r04[3 + q])); ay1[q] = -(r13[3 + q] + (term ? 0.0 : r14[3 + q])); xy0[q] = term ? 0.0 : r04[3 + q]; xy1[q] = term ? 0.0 : r14[3 + q]; }
            yy = r03[5] + r13[5] + (term ? 0.0 : r04[5] + r14[5]) + S[L.dsy + blk];
            if (blk < d) {
                double *Qj = Q + GCS_QN * blk;
                const int o = gcs_uw(blk);
                Qj[0] = aa0[0] + Pu[o]; Qj[1] = aa0[1]; Qj[2] = aa0[2] + Pu[o + 1];
                Qj[3] = aa1[0] + Pu[o + 2]; Qj[4] = aa1[1]; Qj[5] = aa1[2] + Pu[o + 3];
                Qj[6] = ay0[0]; Qj[7] = ay0[1]; Qj[8] = ay1[0]; Qj[9] = ay1[1];
                Qj[10] = yy + Pu[o + 4];
                Qj[11] = xa0[0]; Qj[12] = xa0[1]; Qj[13] = xa0[2]; Qj[14] = xa1[0]; Qj[15] = xa1[1]; Qj[16] = xa1[2];
                Qj[17] = xy0[0]; Qj[18] = xy0[1]; Qj[19] = xy1[0]; Qj[20] = xy1[1];
            } else {    // core block: z_i plays a_i, y_v plays y
                const int Z = GCS_UZ, Y = GCS_UYV, X = GCS_UX;
                C[Z * 10 + Z] = aa0[0]; C[Z * 10 + Z + 1] = C[(Z + 1) * 10 + Z] = aa0[1]; C[(Z + 1) * 10 + Z + 1] = aa0[2];
                C[(Z + 2) * 10 + Z + 2] = aa1[0]; C[(Z + 2) * 10 + Z + 3] = C[(Z + 3) * 10 + Z + 2] = aa1[1]; C[(Z + 3) * 10 + Z + 3] = aa1[2];
                C[Z * 10 + Y] = C[Y * 10 + Z] = ay0[0]; C[(Z + 1) * 10 + Y] = C[Y * 10 + Z + 1] = ay0[1];
                C[(Z + 2) * 10 + Y] = C[Y * 10 + Z + 2] = ay1[0]; C[(Z + 3) * 10 + Y] = C[Y * 10 + Z + 3] = ay1[1];
                C[Y * 10 + Y] = yy;
                C[X * 10 + Z] = C[Z * 10 + X] = xa0[0]; C[X * 10 + Z + 1] = C[(Z + 1) * 10 + X] = xa0[1];
                C[(X + 1) * 10 + Z] = C[Z * 10 + X + 1] = xa0[1]; C[(X + 1) * 10 + Z + 1] = C[(Z + 1) * 10 + X + 1] = xa0[2];
                C[(X + 2) * 10 + Z + 2] = C[(Z + 2) * 10 + X + 2] = xa1[0]; C[(X + 2) * 10 + Z + 3] = C[(Z + 3) * 10 + X + 2] = xa1[1];
                C[(X + 3) * 10 + Z + 2] = C[(Z + 2) * 10 + X + 3] = xa1[1]; C[(X + 3) * 10 + Z + 3] = C[(Z + 3) * 10 + X + 3] = xa1[2];
                C[X * 10 + Y] = C[Y * 10 + X] = xy0[0]; C[(X + 1) * 10 + Y] = C[Y * 10 + X + 1] = xy0[1];
                C[(X + 2) * 10 + Y] = C[Y * 10 + X + 2] = xy1[0]; C[(X + 3) * 10 + Y] = C[Y * 10 + X + 3] = xy1[1];
            }
        }
        GCS_LANE_LOOP(i, 2) {     // x_i x_i collects every C4 | C2 family of point i:  sum D AA'
            const int X = GCS_UX + 2 * i;
            double s0 = 0, s1 = 0, s2 = 0;
            if (!term) for (int blk = 0; blk <= d; ++blk) { const double *r = S + L.eh + GCS_EHP * gcs_slot(blk, i, 1); s0 += r[0]; s1 += r[1]; s2 += r[2]; }
            C[X * 10 + X] = s0; C[X * 10 + X + 1] = s1; C[(X + 1) * 10 + X] = s1; C[(X + 1) * 10 + X + 1] = s2;
        }
        GCS_SYNC();
        if (lane == 0) {
            // second-order cone block  B' W^-2 B  on (t, z1 - z2); a few ulps of its trace keep it PSD
            // W^-2 = (2 wh wh' - J) / beta^2  with  wh = J w   (the inverse of a hyperbolic rotation is the rotation of J w)
            double Wi2[3][3];
            {
                const double wh[3] = {nt.w0, -nt.w1, -nt.w2}, ib2 = 1.0 / (nt.beta * nt.beta);
                for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c)
                    Wi2[a][c] = (2.0 * wh[a] * wh[c] - (a == c ? (a == 0 ? 1.0 : -1.0) : 0.0)) * ib2;
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
        // ---- H_v = N' H_u N (lower triangle) from the structured pieces, then factor ----------
        //   w_j rows/cols:  delta_jk M_j + cs_j cs_k M_* + cz_j cz_k C_zz        (cs = -1 primary, +1 secondary; cz = secondary)
        //   x cols        :  B_j + cs_j B_* + cz_j C_xz (+ cs_j M_* for 's'/'t', where x also feeds z and w_*)
        //   t col         :  cz_j C_tz
        {
            const double *Qs = Q + GCS_QN * jstar;
            double *MS = S + L.MS, *CZ = S + L.CZ;
            GCS_LANE_LOOP(e, 25) { const int a = e / 5, c = e - 5 * a; MS[e] = gcs_qM(Qs, a, c); CZ[e] = C[(GCS_UZ + a) * 10 + GCS_UZ + c]; }
            GCS_SYNC();
            for (int jj = 0; jj < d - 1; ++jj) {
                const int j = jj < jstar ? jj : jj + 1;
                const double csj = prim[j] ? -1.0 : 1.0, czj = prim[j] ? 0.0 : 1.0;
                const double *Qj = Q + GCS_QN * j;
                GCS_LANE_LOOP(e, 25 * (jj + 1)) {
                    const int kk = e / 25, ab = e - 25 * kk, a = ab / 5, c = ab - 5 * a;
                    if (kk == jj && c > a) continue;                       // packed storage: lower triangle only
                    const int k = kk < jstar ? kk : kk + 1;
                    const double csk = prim[k] ? -1.0 : 1.0, czk = prim[k] ? 0.0 : 1.0;
                    double sv = csj * csk * MS[ab] + czj * czk * CZ[ab];
                    if (kk == jj) sv += gcs_qM(Qj, a, c) + (a == c ? 1e-14 : 0.0);
                    H[gcs_tri(5 + 5 * jj + a) + 5 + 5 * kk + c] = sv;
                }
            }
            GCS_LANE_LOOP(e, 25 * (d - 1)) {
                const int r = e / 5, c = e - 5 * r, jj = r / 5, a = r - 5 * jj, j = jj < jstar ? jj : jj + 1;
                const double csj = prim[j] ? -1.0 : 1.0, czj = prim[j] ? 0.0 : 1.0;
                double sv;
                if (c < 4) sv = gcs_qB(Q + GCS_QN * j, c, a) + csj * gcs_qB(Qs, c, a) + czj * C[c * 10 + GCS_UZ + a] + (term ? csj * MS[5 * a + c] : 0.0);
                else sv = czj * C[GCS_UT * 10 + GCS_UZ + a];
                H[gcs_tri(5 + r) + c] = sv;
            }
            GCS_LANE_LOOP(e, 25) {
                const int p = e / 5, q = e - 5 * p;
                if (q <= p) {
                    double sv;
                    if (p < 4) {
                        sv = C[p * 10 + q];
                        if (term) sv += C[p * 10 + GCS_UZ + q] + C[(GCS_UZ + p) * 10 + q] + C[(GCS_UZ + p) * 10 + GCS_UZ + q] + gcs_qM(Qs, p, q) + gcs_qB(Qs, p, q) + gcs_qB(Qs, q, p);
                    } else if (q < 4) sv = C[GCS_UT * 10 + q] + (term ? C[GCS_UT * 10 + GCS_UZ + q] : 0.0);
                    else sv = C[GCS_UT * 10 + GCS_UT];
                    H[gcs_tri(p) + q] = sv + (p == q ? 1e-14 : 0.0);
                }
            }
            GCS_SYNC();
            GCS_LANE_LOOP(q, n) S[L.diag0 + q] = H[gcs_tri(q) + q];
            GCS_SYNC();
        }
        gcs_cholesky(H, d, S + L.diag0, S + L.Linv, lane);

        // ---- predictor:  rhs = -N'(P u + q) -----------------------------------------------------
        GCS_LANE_LOOP(q, nu) ru[q] = -(Pu[q] * u[q] + qu[q]);
        GCS_SYNC();
        gcs_adjoint(ru, dv, d, jstar, prim, term, lane);
        gcs_chol_solve(H, d, S + L.Linv, dv, S + L.ytmp, lane);
        gcs_forward(dv, dua, d, jstar, prim, term, false, lane);
        double tmax = 0.0, dummy = 0.0;
        { GcsRowsArgs ar; ar.mode = 1; ar.sigmu = 0; ar.alpha = 0; gcs_rows(L, S, m, d, term, ar, tmax, dummy, lane); }
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
        { GcsRowsArgs ar; ar.mode = 2; ar.sigmu = sigma * mu; ar.alpha = 0; gcs_rows(L, S, m, d, term, ar, dummy, dummy, lane); }
        gcs_fold(L, S, d, term, ru, -1.0, lane);
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
        gcs_chol_solve(H, d, S + L.Linv, dv, S + L.ytmp, lane);
        gcs_forward(dv, du, d, jstar, prim, term, false, lane);
        { GcsRowsArgs ar; ar.mode = 3; ar.sigmu = sigma * mu; ar.alpha = 0; gcs_rows(L, S, m, d, term, ar, tmax, dummy, lane); }
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
            { GcsRowsArgs ar; ar.mode = 4; ar.sigmu = 0; ar.alpha = alpha; gcs_rows(L, S, m, d, term, ar, mn, sum, lane); }
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
        { GcsRowsArgs ar; ar.mode = 5; ar.sigmu = 0; ar.alpha = alpha; gcs_rows(L, S, m, d, term, ar, dummy, dummy, lane); }
        zq[0] += alpha * dzq[0]; zq[1] += alpha * dzq[1]; zq[2] += alpha * dzq[2];
        GCS_SYNC();
        gcs_forward(v, u, d, jstar, prim, term, true, lane);
    }
    if (in.ws != 0) {   // warm-start record for the next ADMM iteration (only a converged primal-dual pair is reused)
        const int nrw = 4 * (L.dcap + 1) * L.mcap;
        double *wv = in.ws + 1, *wz = in.ws + 1 + L.ncap, *wy = wz + nrw, *wq = wy + L.dcap + 1;
        if (res.status == 0) {
            GCS_LANE_LOOP(q, n) wv[q] = v[q];
            { const int NS = 4 * (L.dcap + 1), ns = 4 * (d + 1);
              GCS_LANE_LOOP(q, ns * m) { const int k = q / ns, idx = k * NS + (q - k * ns); wz[idx] = S[L.zr + idx]; } }
            GCS_LANE_LOOP(j, d + 1) wy[j] = S[L.zy + j];
            if (lane == 0) { wq[0] = zq[0]; wq[1] = zq[1]; wq[2] = zq[2]; in.ws[0] = 1.0; }
        } else if (lane == 0) in.ws[0] = 0.0;
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
