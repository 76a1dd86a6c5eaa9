// K1 wrapper: x-update of one vertex (reference admm_solver_v3.py:352-540) on the flat half-edge layout.
// Gathers the consensus targets z_e + mu_h of the vertex's half-edges, runs the warp-level interior
// point solve (vertex_ipm.cuh) and scatters the consensus copies xc_h plus x_v, z_v, y_v.
#pragma once
#include "vertex_ipm.cuh"

#define GCS_HE_OUT 1   // owner is the edge's tail
#define GCS_HE_ZERO 2  // flow forced to 0 (presolve): in-edges of 's', out-edges of 't', edges of flow-less vertices
#define GCS_VT_GENERIC 0
#define GCS_VT_SOURCE 1
#define GCS_VT_TARGET 2
#define GCS_VT_DEAD 3

struct GcsGraphView {
    int nV, nE;
    const int *poly_off; const double *polyA, *polyb;
    const int *he_off, *he_edge; const unsigned char *he_flags;
    const unsigned char *vtype; const double *cent;
};
struct GcsStateView {
    double *xc;          // [H][5]  consensus copies, edge-canonical order (z_u[:2], z_w[:2], y)
    const double *mu;    // [H][5]  scaled duals
    const double *z;     // [E][5]  edge variables
    double *x_v, *z_v, *y_v;   // [nV][4], [nV][4], [nV]
    double *ws;                // [nV][gcs_ws_stride] warm-start records, or null
    double theta;
    double zero_tol;           // targets with max-norm <= zero_tol are treated as exactly zero
};

// returns the interior-point iteration count (lane-uniform); *status receives the solve status
GCS_DEV int gcs_vertex_update(const GcsGraphView &G, const GcsStateView &St, int v, double rho, double mu_scale,
                              double tol, int max_iter, const GcsScratchLayout &L, double *S, int lane, int *status) {
    const int h0 = G.he_off[v], h1 = G.he_off[v + 1];
    const int type = G.vtype[v];
    *status = 0;
    // forced-zero half-edges: own copy and flow are 0; an incoming edge's "other copy" first point is
    // unconstrained in the vertex program and therefore sits exactly on its target
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
    if (type == GCS_VT_DEAD) {
        if (lane == 0) {
            for (int k = 0; k < 4; ++k) { St.z_v[4 * (size_t)v + k] = 0.0; St.x_v[4 * (size_t)v + k] = G.cent[2 * (size_t)v + (k & 1)]; }
            St.y_v[v] = 0.0;
        }
        return 0;
    }
    int *out = (int *)(S + L.ints), *hid = out + 2 * L.dcap;
    int d = 0;
    for (int h = h0; h < h1; ++h) {        // lane-uniform compaction of the live half-edges
        const int f = G.he_flags[h];
        if (f & GCS_HE_ZERO) continue;
        if (lane == 0) { out[d] = f & GCS_HE_OUT; hid[d] = h; }
        d++;
    }
    const int p0 = G.poly_off[v], m = G.poly_off[v + 1] - p0;
    GCS_LANE_LOOP(k, m) { S[L.A + 2 * k] = G.polyA[2 * (size_t)(p0 + k)]; S[L.A + 2 * k + 1] = G.polyA[2 * (size_t)(p0 + k) + 1]; S[L.b + k] = G.polyb[p0 + k]; }
    GCS_SYNC();
    GCS_LANE_LOOP(q, 5 * d) {
        const int j = q / 5, c = q - 5 * j, h = hid[j], e = G.he_edge[h];
        S[L.tgt + q] = St.z[5 * (size_t)e + c] + mu_scale * St.mu[5 * (size_t)h + c];
    }
    GCS_SYNC();
    if (type == GCS_VT_GENERIC) {
        // all consensus targets (numerically) zero — untouched or long-decayed region: for zero targets the
        // program's optimum is the origin (cost t + eps y + rho/2 |.|^2 >= 0, attained at a = 0, y = 0, t = 0),
        // and the prox map is non-expansive, so |solution| <= |targets| <= zero_tol (1e-12 by default): no solve needed
        double nz = 0.0;
        GCS_LANE_LOOP(q, 5 * d) nz = fmax(nz, fabs(S[L.tgt + q]));
        nz = gcs_warp_max(nz);
        if (nz <= St.zero_tol) {
            GCS_LANE_LOOP(q, 5 * d) { const int j = q / 5; St.xc[5 * (size_t)hid[j] + (q - 5 * j)] = 0.0; }
            if (lane == 0) {
                for (int k = 0; k < 4; ++k) { St.z_v[4 * (size_t)v + k] = 0.0; St.x_v[4 * (size_t)v + k] = G.cent[2 * (size_t)v + (k & 1)]; }
                St.y_v[v] = 0.0;
                if (St.ws) St.ws[(size_t)v * gcs_ws_stride(L)] = 0.0;
            }
            GCS_SYNC();
            *status = -1;     // skipped
            return 0;
        }
    }
    GcsVertexIn in;
    in.m = m; in.d = d; in.type = type; in.cx = G.cent[2 * (size_t)v]; in.cy = G.cent[2 * (size_t)v + 1];
    in.rho = rho; in.tol = tol; in.max_iter = max_iter;
    in.ws = St.ws ? St.ws + (size_t)v * gcs_ws_stride(L) : 0; in.theta = St.theta;
    GcsVertexOut r = gcs_vertex_solve(L, S, in, lane);
    *status = r.status;
    const double *u = S + L.u, *tgt = S + L.tgt;
    GCS_LANE_LOOP(j, d) {     // scatter (:492-522), edge-canonical order
        const double *w = u + gcs_uw(j), *t = tgt + 5 * j;
        double *x = St.xc + 5 * (size_t)hid[j];
        if (out[j]) { x[0] = w[0]; x[1] = w[1]; x[2] = w[2]; x[3] = w[3]; }     // own first point | other's first point == own second point (C5)
        else        { x[0] = t[0]; x[1] = t[1]; x[2] = w[0]; x[3] = w[1]; }     // other's first point is free -> its target | own first point
        x[4] = w[4];
    }
    if (lane == 0) {
        for (int k = 0; k < 4; ++k) { St.x_v[4 * (size_t)v + k] = u[GCS_UX + k]; St.z_v[4 * (size_t)v + k] = u[GCS_UZ + k]; }
        St.y_v[v] = u[GCS_UYV];
    }
    GCS_SYNC();
    return r.iters;
}
