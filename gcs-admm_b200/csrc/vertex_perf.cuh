// K1 in `perf` mode — the x-update of the full-vertex-split ADMM done INEXACTLY by K warm-started iterations
// of an operator-splitting scheme whose steps are all closed-form (north_star: "fixed-iteration inner
// projection / primal-dual scheme").  One warp per vertex, state in shared memory, same null-space
// parametrisation u = N v + up as the exact kernel (vertex_ipm.cuh).
//
// Every inequality of the vertex program (reference admm_solver_v3.py:416-440) says that a (point, flow) pair
// lies in the perspective cone of the vertex's polygon,  K_P = {(p, h) in R^3 : A p <= h b}:
//     C3: (a_i, y_e)        C4: (x_i - a_i, 1 - y_e)        C1: (z_i, y_v)        C2: (x_i - z_i, 1 - y_v)
// and the pairs are 0/+-1 linear images  c = M u + m0  of the variables.  Splitting  c in prod K_P  from u:
//     v-step : minimise  rho/2 |S u - T|^2 + eps'y + sigma/2 |M u + m0 - c + lam|^2   over v   (u = N v + up)
//              -> v <- v - (1/rho) K1^-1 G(v),  K1 = N'(S'S + kappa M'M)N  depends only on the vertex CLASS
//              (type, #in, #out) — not on the polygon —, so its inverse is a table shared by all vertices;
//     c-step : c <- Proj_{K_P}(alpha (M u + m0) + (1 - alpha) c + lam)   exact 3-D cone projection, O(#polygon vertices);
//              the path-length term |z_1 - z_2| (:380-384) is a block soft-threshold with threshold 1/sigma;
//     lam    : lam <- lam + alpha (M u + m0) + (1 - alpha) c_old - c_new.
// sigma = kappa * rho, so the rho-adaptation rescale of the ADMM duals (:705/:708) applies to lam as well.
// (c, lam) persist per vertex in HBM between ADMM iterations (warm start).
//
// The fixed point of the outer ADMM is unchanged (an exact minimiser of the vertex program is a fixed point of
// the inner iteration); the trajectory is not the reference's, so this mode is validated at convergence against
// the classic relaxation optimum (tests/test_gpu_perf.py, tools/prototypes/inner_first_order.py), not per iteration.
#pragma once
#include "vertex_update.cuh"
#define GCS_CONE_REC 12   // doubles per polygon vertex in the cone table
// the per-vertex (c, lam) records are streamed once per iteration: evict-first loads / stores keep L2 for xc, z, mu and the tables
#if defined(__CUDA_ARCH__)
#define GCS_LD_STREAM(p) __ldcs(p)
#define GCS_ST_STREAM(p, x) __stcs((p), (x))
#else
#define GCS_LD_STREAM(p) (*(p))
#define GCS_ST_STREAM(p, x) (*(p) = (x))
#endif

struct GcsPerfTables {
    const int *vclass;         // [nV] class id (-1: vertex not solved here: dead)
    const int *class_koff;     // [ncls] offset of the class's K1^-1 (n x n, row-major) in kinv
    const double *kinv;
    const int *cone_off;       // [nV+1] polygon vertices of vertex v: cone_off[v]..cone_off[v+1]
    const double *cone;        // GCS_CONE_REC doubles per polygon vertex (see gcs_cone_project)
    double *state;             // [nV][state_stride]: c then lam, (3 * 4 (dcap + 1) + 2) doubles each
    int state_stride;
    int inner_iters;           // K
    double alpha, kappa;
};

// scratch (doubles) of one warp in perf mode
struct GcsPerfLayout { int dcap, kcap, ncap, nucap, npair, u, gu, v, gv, pv, c, lam, w, cone, tgt, ints, total; };
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline GcsPerfLayout gcs_perf_layout(int dcap, int kcap) {
    GcsPerfLayout L;
    if (dcap < 1) dcap = 1;
    if (kcap < 3) kcap = 3;
    L.dcap = dcap; L.kcap = kcap; L.ncap = 5 * dcap; L.nucap = GCS_NCORE + 5 * dcap; L.npair = 4 * (dcap + 1);
    int o = 0;
    L.u = o; o += L.nucap; L.gu = o; o += L.nucap;
    L.v = o; o += L.ncap; L.gv = o; o += L.ncap;
    const int np3 = 3 * L.npair + 2;
    L.pv = o; o += np3; L.c = o; o += np3; L.lam = o; o += np3; L.w = o; o += np3;
    L.cone = o; o += GCS_CONE_REC * kcap;
    L.tgt = o; o += 5 * dcap;
    L.ints = o; o += (3 * dcap + 1) / 2 + 1;
    L.total = o;
    return L;
}
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int gcs_perf_state_stride(int dcap) { return 2 * (3 * 4 * (dcap + 1) + 2); }

// exact projection of c onto the cone spanned by the rays r_k = (V_k, 1), k = 0..nv-1 (counter-clockwise).
// record k: V_k (2) | unit outward normal n_k of the face between rays k and k+1 (3) | 1/|r_k|^2 | sector normals ma, mb (3 + 3)
// The projection is c itself, or lies on a face (inside its sector), on a ray, or is the apex: take the nearest candidate.
// Candidates are coded  -1 apex | 2k ray k | 2k+1 face k  and scanned in that order with a strict "<", so the result does
// not depend on how the scan is split over lanes (ties go to the smallest code).
GCS_DEV void gcs_cone_scan(const double *cone, int nv, int first, int step, double c0, double c1, double c2, double &bd, int &code, bool &inside) {
    for (int k = first; k < nv; k += step) {
        const double *ck = cone + GCS_CONE_REC * k;
        const double rx = ck[0], ry = ck[1];
        const double tau = (c0 * rx + c1 * ry + c2) * ck[5];                  // ray k
        if (tau > 0.0) {
            const double e0 = c0 - tau * rx, e1 = c1 - tau * ry, e2 = c2 - tau;
            const double dd = e0 * e0 + e1 * e1 + e2 * e2;
            if (dd < bd) { bd = dd; code = 2 * k; }
        }
        const double nx = ck[2], ny = ck[3], nh = ck[4];
        const double dist = nx * c0 + ny * c1 + nh * c2;
        if (dist > 0.0) {                                                      // outside face k
            inside = false;
            const double p0 = c0 - dist * nx, p1 = c1 - dist * ny, p2 = c2 - dist * nh;
            if (p0 * ck[6] + p1 * ck[7] + p2 * ck[8] >= 0.0 && p0 * ck[9] + p1 * ck[10] + p2 * ck[11] >= 0.0) {
                const double dd = dist * dist;
                if (dd < bd) { bd = dd; code = 2 * k + 1; }
            }
        }
    }
}
GCS_DEV void gcs_cone_point(const double *cone, int code, bool inside, double c0, double c1, double c2, double &q0, double &q1, double &q2) {
    if (inside) { q0 = c0; q1 = c1; q2 = c2; return; }
    q0 = 0.0; q1 = 0.0; q2 = 0.0;
    if (code < 0) return;
    const double *ck = cone + GCS_CONE_REC * (code >> 1);
    if (code & 1) {
        const double nx = ck[2], ny = ck[3], nh = ck[4];
        const double dist = nx * c0 + ny * c1 + nh * c2;
        q0 = c0 - dist * nx; q1 = c1 - dist * ny; q2 = c2 - dist * nh;
    } else {
        const double rx = ck[0], ry = ck[1];
        const double tau = (c0 * rx + c1 * ry + c2) * ck[5];
        q0 = tau * rx; q1 = tau * ry; q2 = tau;
    }
}
GCS_DEV void gcs_cone_project(const double *cone, int nv, double c0, double c1, double c2, double &q0, double &q1, double &q2) {
    double bd = c0 * c0 + c1 * c1 + c2 * c2;     // the apex
    int code = -1;
    bool inside = true;
    gcs_cone_scan(cone, nv, 0, 1, c0, c1, c2, bd, code, inside);
    gcs_cone_point(cone, code, inside, c0, c1, c2, q0, q1, q2);
}

// pair values  pv = M u + m0  (3 per family slot, then the 2 entries of z_1 - z_2)
GCS_DEV void gcs_pair_values(const double *u, double *pv, int d, bool term, int npair_cap, int lane) {
    const int nitems = term ? 2 * (d + 1) : 4 * (d + 1);
    GCS_LANE_LOOP(it, nitems) {
        int fam, i, blk;
        if (term) { fam = 0; i = it & 1; blk = it >> 1; } else { fam = it & 1; i = (it >> 1) & 1; blk = it >> 2; }
        const int po = (blk < d ? gcs_uw(blk) : GCS_UZ) + 2 * i, yo = blk < d ? gcs_uw(blk) + 4 : GCS_UYV, xo = GCS_UX + 2 * i;
        double *p = pv + 3 * gcs_slot(blk, i, fam);
        p[0] = fam ? u[xo] - u[po] : u[po];
        p[1] = fam ? u[xo + 1] - u[po + 1] : u[po + 1];
        p[2] = fam ? 1.0 - u[yo] : u[yo];
    }
    if (lane == 0) {
        pv[3 * npair_cap] = u[GCS_UZ] - u[GCS_UZ + 2];
        pv[3 * npair_cap + 1] = u[GCS_UZ + 1] - u[GCS_UZ + 3];
    }
    GCS_SYNC();
}

// gu = M' w  (w: 3 per slot + 2)
GCS_DEV void gcs_pair_adjoint(const double *w, double *gu, int d, bool term, int npair_cap, int lane) {
    GCS_LANE_LOOP(q, 4 * (d + 1)) {          // point slots a_i / z_i
        const int blk = q >> 2, i = (q >> 1) & 1, c = q & 1;
        const double *w3 = w + 3 * gcs_slot(blk, i, 0);
        gu[(blk < d ? gcs_uw(blk) : GCS_UZ) + 2 * i + c] = w3[c] - (term ? 0.0 : w3[3 + c]);
    }
    GCS_LANE_LOOP(blk, d + 1) {              // flow slots y / y_v
        const double *w0 = w + 3 * gcs_slot(blk, 0, 0);
        double s = w0[2] + w0[6 + 2];
        if (!term) s -= w0[3 + 2] + w0[9 + 2];
        gu[blk < d ? gcs_uw(blk) + 4 : GCS_UYV] = s;
    }
    GCS_LANE_LOOP(q, 4) {                    // x_i collects every C4 | C2 pair of point i
        const int i = q >> 1, c = q & 1;
        double s = 0.0;
        if (!term) for (int blk = 0; blk <= d; ++blk) s += w[3 * gcs_slot(blk, i, 1) + c];
        gu[GCS_UX + q] = s;
    }
    GCS_SYNC();
    if (lane == 0) {
        gu[GCS_UT] = 0.0;
        const double n0 = w[3 * npair_cap], n1 = w[3 * npair_cap + 1];
        gu[GCS_UZ] += n0; gu[GCS_UZ + 1] += n1; gu[GCS_UZ + 2] -= n0; gu[GCS_UZ + 3] -= n1;
    }
    GCS_SYNC();
}

// x-update of one vertex in perf mode.  Returns 1 (the vertex did K inner iterations) or 0 (no program).
GCS_DEV int gcs_vertex_update_perf(const GcsGraphView &G, const GcsStateView &St, const GcsPerfTables &T, int v, double rho,
                                   double mu_scale, const GcsPerfLayout &L, double *S, int lane) {
    const int h0 = G.he_off[v], h1 = G.he_off[v + 1];
    const int type = G.vtype[v];
    // Loads are issued in two dependent stages only: (flags, edge ids, cone records, stored state) -> (edge variables, duals).
#ifndef GCS_EMULATE
    const bool fast = h1 - h0 <= 32;           // one lane per half-edge keeps its flags / edge id in registers
    int f = GCS_HE_ZERO, e_l = 0;
    if (fast && lane < h1 - h0) { f = G.he_flags[h0 + lane]; e_l = G.he_edge[h0 + lane]; }
#else
    const bool fast = false;
    const int f = 0, e_l = 0;
#endif
    const int np3 = 3 * L.npair + 2;
    const int c0 = T.cone_off[v], nv = T.cone_off[v + 1] - c0;
    double *st = T.state + (size_t)v * T.state_stride;
    if (type != GCS_VT_DEAD) {
        GCS_LANE_LOOP(q, GCS_CONE_REC * nv) S[L.cone + q] = T.cone[GCS_CONE_REC * (size_t)c0 + q];
        GCS_LANE_LOOP(q, np3) { S[L.c + q] = GCS_LD_STREAM(st + q); S[L.lam + q] = mu_scale * GCS_LD_STREAM(st + np3 + q); }    // sigma = kappa rho: lam rescales with mu
    }
    if (fast) {                                 // forced-zero half-edges, as in the exact kernel
        if (lane < h1 - h0 && (f & GCS_HE_ZERO)) {
            const int h = h0 + lane;
            double *x = St.xc + 5 * (size_t)h;
            double x0 = 0.0, x1 = 0.0;
            if (!(f & GCS_HE_OUT)) {
                x0 = St.z[5 * (size_t)e_l] + mu_scale * St.mu[5 * (size_t)h];
                x1 = St.z[5 * (size_t)e_l + 1] + mu_scale * St.mu[5 * (size_t)h + 1];
            }
            x[0] = x0; x[1] = x1; x[2] = 0.0; x[3] = 0.0; x[4] = 0.0;
        }
    } else {
        GCS_LANE_LOOP(i, h1 - h0) {
            const int h = h0 + i;
            if (G.he_flags[h] & GCS_HE_ZERO) {
                double *x = St.xc + 5 * (size_t)h;
                double x0 = 0.0, x1 = 0.0;
                if (!(G.he_flags[h] & GCS_HE_OUT)) {
                    const int e = G.he_edge[h];
                    x0 = St.z[5 * (size_t)e] + mu_scale * St.mu[5 * (size_t)h];
                    x1 = St.z[5 * (size_t)e + 1] + mu_scale * St.mu[5 * (size_t)h + 1];
                }
                x[0] = x0; x[1] = x1; x[2] = 0.0; x[3] = 0.0; x[4] = 0.0;
            }
        }
    }
    if (type == GCS_VT_DEAD) {
        if (lane == 0) {
            for (int k = 0; k < 4; ++k) { St.z_v[4 * (size_t)v + k] = 0.0; St.x_v[4 * (size_t)v + k] = G.cent[2 * (size_t)v + (k & 1)]; }
            St.y_v[v] = 0.0;
        }
        return 0;
    }
    int *out = (int *)(S + L.ints), *prim = out + L.dcap, *hid = out + 2 * L.dcap;
    int d = 0, jstar = -1;
#ifndef GCS_EMULATE
    if (fast) {                // live list by ballot + prefix popcount instead of a serial scan; each live lane gathers its own targets
        const bool live = !(f & GCS_HE_ZERO);
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (live) {
            const int k = __popc(m & ((1u << lane) - 1u)), h = h0 + lane;
            out[k] = f & GCS_HE_OUT; hid[k] = h; prim[k] = (type == GCS_VT_TARGET) ? 1 : (f & GCS_HE_OUT);
            const double *zz = St.z + 5 * (size_t)e_l, *mm = St.mu + 5 * (size_t)h;
            double *t = S + L.tgt + 5 * k;
            for (int c = 0; c < 5; ++c) t[c] = zz[c] + mu_scale * mm[c];
        }
        d = __popc(m);
        const unsigned pm = __ballot_sync(0xffffffffu, live && ((type == GCS_VT_TARGET) || (f & GCS_HE_OUT)));
        if (pm) jstar = __popc(m & ((1u << (31 - __clz(pm))) - 1u));      // the last primary block is the dependent one
    } else
#endif
    {
        for (int h = h0; h < h1; ++h) {
            const int fl = G.he_flags[h];
            if (fl & GCS_HE_ZERO) continue;
            if (lane == 0) { out[d] = fl & GCS_HE_OUT; hid[d] = h; prim[d] = (type == GCS_VT_TARGET) ? 1 : (fl & GCS_HE_OUT); }
            d++;
        }
        GCS_SYNC();
        for (int j = 0; j < d; ++j) if (prim[j]) jstar = j;
        GCS_LANE_LOOP(q, 5 * d) {
            const int j = q / 5, c = q - 5 * j, h = hid[j], e = G.he_edge[h];
            S[L.tgt + q] = St.z[5 * (size_t)e + c] + mu_scale * St.mu[5 * (size_t)h + c];
        }
    }
    const bool term = type != GCS_VT_GENERIC;
    const int n = 5 * d, nu = GCS_NCORE + 5 * d;
    GCS_SYNC();
    double *u = S + L.u, *gu = S + L.gu, *vv = S + L.v, *gv = S + L.gv, *pv = S + L.pv, *cc = S + L.c, *lam = S + L.lam, *w = S + L.w;
    const double *tgt = S + L.tgt;
    const double sigma = T.kappa * rho, alpha = T.alpha;
    const double *Kinv = T.kinv + T.class_koff[T.vclass[v]];
    // start: v with  N v + up  closest to the stored pair copies is not needed — the v-step is an exact solve of a
    // quadratic, so any starting v gives the same result; start from 0
    GCS_LANE_LOOP(q, n) vv[q] = 0.0;
    if (!term) {
        // generic vertex: u(0) = up = 0, so the pair values are the constants  (a_i, y) = 0,  (x_i - a_i, 1 - y) = (0, 0, 1)
        GCS_LANE_LOOP(q, nu) u[q] = 0.0;
        GCS_LANE_LOOP(q, np3) pv[q] = 0.0;
        GCS_SYNC();
        GCS_LANE_LOOP(q, 2 * (d + 1)) pv[3 * gcs_slot(q >> 1, q & 1, 1) + 2] = 1.0;
        GCS_SYNC();
    } else {
        GCS_SYNC();
        gcs_forward(vv, u, d, jstar, prim, term, true, lane);
        gcs_pair_values(u, pv, d, term, L.npair, lane);
    }
    for (int it = 0; it < T.inner_iters; ++it) {
        // G_u = rho S'(S u - T) + eps e_y + sigma M'(pv - c + lam)
        GCS_LANE_LOOP(q, np3) w[q] = pv[q] - cc[q] + lam[q];
        GCS_SYNC();
        gcs_pair_adjoint(w, gu, d, term, L.npair, lane);
        GCS_LANE_LOOP(q, nu) {
            double g = sigma * gu[q];
            if (q >= GCS_NCORE) {
                const int j = (q - GCS_NCORE) / 5, c = q - GCS_NCORE - 5 * j;
                const double *t = tgt + 5 * j;
                if (out[j]) { if (c < 4) g += rho * (u[q] - t[c]); }
                else if (c < 2) g += rho * (u[q] - t[2 + c]);
                if (c == 4) g += rho * (u[q] - t[4]) + GCS_EDGE_PENALTY;
            }
            gu[q] = g;
        }
        GCS_SYNC();
        gcs_adjoint(gu, gv, d, jstar, prim, term, lane);
        // v <- v - (1/rho) K1^-1 G_v      (dense n x n table of the vertex class)
        const double irho = 1.0 / rho;
#ifdef GCS_EMULATE
        const int mv_full = n;
#else
        const int mv_full = n & ~31;                  // whole rounds of 32 rows: one lane per row
#endif
        GCS_LANE_LOOP(r, mv_full) {
            // K1^-1 is symmetric: walk column r (= row r) so that the lanes of a warp read consecutive doubles of
            // row k — coalesced, L1-resident table shared by every vertex of the class
            const double *kc = Kinv + r;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = 0;
            for (; k + 7 < n; k += 8) {                   // 8 independent table loads in flight per lane
                const double a0 = kc[(size_t)k * n], a1 = kc[(size_t)(k + 1) * n], a2 = kc[(size_t)(k + 2) * n], a3 = kc[(size_t)(k + 3) * n];
                const double a4 = kc[(size_t)(k + 4) * n], a5 = kc[(size_t)(k + 5) * n], a6 = kc[(size_t)(k + 6) * n], a7 = kc[(size_t)(k + 7) * n];
                s0 += a0 * gv[k]; s1 += a1 * gv[k + 1]; s2 += a2 * gv[k + 2]; s3 += a3 * gv[k + 3];
                s0 += a4 * gv[k + 4]; s1 += a5 * gv[k + 5]; s2 += a6 * gv[k + 6]; s3 += a7 * gv[k + 7];
            }
            for (; k < n; ++k) s0 += kc[(size_t)k * n] * gv[k];
            w[r] = vv[r] - irho * ((s0 + s1) + (s2 + s3));            // w doubles as the new v until every lane has read gv / vv
        }
#ifndef GCS_EMULATE
        if (n > mv_full) {       // the last R < 32 rows: lp = 2^k lanes per row split the columns, partial sums reduced by shuffles
            const int R = n - mv_full;
            int lp = 1, lg = 0;
            while (2 * lp * R <= 32) { lp *= 2; ++lg; }
            const int grp = lane >> lg, sub = lane & (lp - 1), r = mv_full + (grp < R ? grp : 0);
            const double *kc = Kinv + r;
            double s0 = 0.0, s1 = 0.0;
            int k = sub;
            for (; k + lp < n; k += 2 * lp) {
                s0 += kc[(size_t)k * n] * gv[k];
                s1 += kc[(size_t)(k + lp) * n] * gv[k + lp];
            }
            if (k < n) s0 += kc[(size_t)k * n] * gv[k];
            double sum = s0 + s1;
            for (int m = 1; m < lp; m <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
            if (grp < R && sub == 0) w[r] = vv[r] - irho * sum;
        }
#endif
        GCS_SYNC();
        GCS_LANE_LOOP(r, n) vv[r] = (r == 4) ? 0.0 : w[r];     // the epigraph variable t is unused in this mode
        GCS_SYNC();
        gcs_forward(vv, u, d, jstar, prim, term, true, lane);
        gcs_pair_values(u, pv, d, term, L.npair, lane);
        // c-step and dual step: item idx < nitems is a (point, flow) pair, item nitems is |z_1 - z_2|
        const int nitems = term ? 2 * (d + 1) : 4 * (d + 1);
        auto pair_slot = [&](int idx) {
            int fam, i, blk;
            if (term) { fam = 0; i = idx & 1; blk = idx >> 1; } else { fam = idx & 1; i = (idx >> 1) & 1; blk = idx >> 2; }
            return 3 * gcs_slot(blk, i, fam);
        };
        auto norm_item = [&]() {           // block soft-threshold, threshold 1 / sigma
            const int o = 3 * L.npair;
            const double r0 = alpha * pv[o] + (1.0 - alpha) * cc[o], r1 = alpha * pv[o + 1] + (1.0 - alpha) * cc[o + 1];
            const double a0 = r0 + lam[o], a1 = r1 + lam[o + 1], nrm = hypot(a0, a1);
            const double sc = nrm > 0.0 ? fmax(0.0, 1.0 - 1.0 / (sigma * nrm)) : 0.0;
            const double q0 = sc * a0, q1 = sc * a1;
            lam[o] += r0 - q0; lam[o + 1] += r1 - q1;
            cc[o] = q0; cc[o + 1] = q1;
        };
#ifdef GCS_EMULATE
        const int full_end = nitems + 1;
#else
        const int full_end = (nitems + 1) & ~31;      // whole rounds of 32 items: one lane per item
#endif
        GCS_LANE_LOOP(idx, full_end) {
            if (idx < nitems) {
                const int o = pair_slot(idx);
                const double r0 = alpha * pv[o] + (1.0 - alpha) * cc[o], r1 = alpha * pv[o + 1] + (1.0 - alpha) * cc[o + 1],
                             r2 = alpha * pv[o + 2] + (1.0 - alpha) * cc[o + 2];
                double q0, q1, q2;
                gcs_cone_project(S + L.cone, nv, r0 + lam[o], r1 + lam[o + 1], r2 + lam[o + 2], q0, q1, q2);
                lam[o] += r0 - q0; lam[o + 1] += r1 - q1; lam[o + 2] += r2 - q2;
                cc[o] = q0; cc[o + 1] = q1; cc[o + 2] = q2;
            } else norm_item();
        }
#ifndef GCS_EMULATE
        {   // the last R < 32 items: lp = 2^k lanes per item share the scan of its faces, then reduce (value, code) by shuffles
            const int R = nitems + 1 - full_end;
            if (R > 0) {
                int lp = 1, lg = 0;
                while (2 * lp * R <= 32) { lp *= 2; ++lg; }
                const int grp = lane >> lg, sub = lane & (lp - 1), idx = full_end + grp;
                const bool pair = grp < R && idx < nitems;
                const int o = pair ? pair_slot(idx) : 0;
                double r0 = 0.0, r1 = 0.0, r2 = 0.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;
                if (pair) {
                    r0 = alpha * pv[o] + (1.0 - alpha) * cc[o]; r1 = alpha * pv[o + 1] + (1.0 - alpha) * cc[o + 1];
                    r2 = alpha * pv[o + 2] + (1.0 - alpha) * cc[o + 2];
                    c0 = r0 + lam[o]; c1 = r1 + lam[o + 1]; c2 = r2 + lam[o + 2];
                }
                double bd = c0 * c0 + c1 * c1 + c2 * c2;
                int code = -1;
                bool inside = true;
                gcs_cone_scan(S + L.cone, pair ? nv : 0, sub, lp, c0, c1, c2, bd, code, inside);
                int in_i = inside ? 1 : 0;
                for (int m = 1; m < lp; m <<= 1) {
                    const double obd = __shfl_xor_sync(0xffffffffu, bd, m);
                    const int ocode = __shfl_xor_sync(0xffffffffu, code, m);
                    in_i &= __shfl_xor_sync(0xffffffffu, in_i, m);
                    if (obd < bd || (obd == bd && ocode < code)) { bd = obd; code = ocode; }
                }
                __syncwarp();
                if (pair && sub == 0) {
                    double q0, q1, q2;
                    gcs_cone_point(S + L.cone, code, in_i != 0, c0, c1, c2, q0, q1, q2);
                    lam[o] += r0 - q0; lam[o + 1] += r1 - q1; lam[o + 2] += r2 - q2;
                    cc[o] = q0; cc[o + 1] = q1; cc[o + 2] = q2;
                }
                if (grp < R && idx == nitems && sub == 0) norm_item();
            }
        }
#endif
        GCS_SYNC();
    }
    GCS_LANE_LOOP(q, np3) { GCS_ST_STREAM(st + q, cc[q]); GCS_ST_STREAM(st + np3 + q, lam[q]); }
    GCS_LANE_LOOP(j, d) {     // scatter, edge-canonical order (same as the exact kernel)
        const double *wj = u + gcs_uw(j), *t = tgt + 5 * j;
        double *x = St.xc + 5 * (size_t)hid[j];
        if (out[j]) { x[0] = wj[0]; x[1] = wj[1]; x[2] = wj[2]; x[3] = wj[3]; }
        else        { x[0] = t[0]; x[1] = t[1]; x[2] = wj[0]; x[3] = wj[1]; }
        x[4] = wj[4];
    }
    if (lane == 0) {
        for (int k = 0; k < 4; ++k) { St.x_v[4 * (size_t)v + k] = u[GCS_UX + k]; St.z_v[4 * (size_t)v + k] = u[GCS_UZ + k]; }
        St.y_v[v] = u[GCS_UYV];
    }
    GCS_SYNC();
    return 1;
}
