// Stacking-stability approximation of the PCT environment, restated for one CUDA lane per placement.
//
// Reference semantics (paths relative to the reference repo, D: = pct_envs/PctDiscrete0/):
//   Box.calculated_impact / calculated_impact_virtual   D:space.py:73-267 (recursive, dict-ordered loads)
//   Box.calculate_new_com                               D:space.py:51-71
//   ConvexHull / Line2D / point_in_polygen              D:convex_hull.py:4-112
//   Space.scale_down                                    D:space.py:341-345
//   np.linalg.lstsq for >=3 supports                    D:space.py:137-152, 234-250
//
// B200 restatement (not a translation):
//   * no objects, no recursion: an explicit DFS over the support DAG with a small per-lane frame stack;
//   * supports, contact rectangles, hulls are recomputed from the packed box records in shared memory
//     instead of being stored per box;
//   * the only persistent stability state is the per-env "load edge" pool in HBM
//     (upper box, lower box, load centre xyz, load mass); dict insertion order of the reference's
//     `up_edges` == pool order, because edges are created in placement order and updated in place;
//   * a virtual (feasibility) check never writes: the virtual load of the path member directly above
//     is carried in the frame, and the real edge of that path member is skipped (== the `involved`
//     gating of D:space.py:55-63);
//   * all FP64 arithmetic is one IEEE operation per source operator (file is compiled with
//     -fmad=false); fma() appears only where it reproduces OpenBLAS ddot on 2-vectors;
//   * the step kernel is instruction-fetch bound (ncu: stalled_no_instruction), so this file is written
//     for SMALL CODE: one runtime-flag routine for real and virtual checks, FP64 division / sqrt and the
//     helpers are shared __noinline__ functions, loops are not unrolled.
#pragma once
#include <cstdint>
#include <math.h>
#include <math_constants.h>
#include "pct_kernels.h"

namespace pct {

__device__ __noinline__ double ddiv(double a, double b) { return a / b; }
__device__ __noinline__ double dsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ double dot2(double u0, double u1, double v0, double v1) { return fma(u1, v1, u0 * v0); }

__device__ __forceinline__ double slope_of(double ax, double ay, double bx, double by) {
    if (bx != ax) return ddiv(by - ay, bx - ax);
    return (by - ay) * CUDART_INF;  // 0*inf = nan like the reference (convex_hull.py:14)
}
__device__ __forceinline__ int orient_of(double s1, double s2) {
    if (fabs(s1) == CUDART_INF && fabs(s2) == CUDART_INF) return 0;
    double d = s2 - s1;
    if (d > 0) return -1;
    if (d == 0) return 0;
    return 1;
}

// ConvexHull (convex_hull.py:39-95) on n points already perturbed (x += y*1e-6).
// Writes the hull as indices: lower chain (minus last) then upper chain (minus last) into `hl` (returns count).
__device__ __noinline__ int hull_indices(const double *px, const double *py, int n, uint8_t *order, uint8_t *hl, uint8_t *hu) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {  // stable insertion sort by x (sorted(key=x[0]), :34-37)
        const double kx = px[i];
        int j = i - 1;
#pragma unroll 1
        while (j >= 0 && px[order[j]] > kx) { order[j + 1] = order[j]; j--; }
        order[j + 1] = (uint8_t)i;
    }
    int m = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        uint8_t *H = pass == 0 ? hl : hu;
        int nh = 0;
#pragma unroll 1
        for (int k = 0; k < n; k++) {
            const int p = order[pass == 0 ? k : n - 1 - k];
            const double qx = px[p], qy = py[p];
#pragma unroll 1
            while (nh >= 2) {
                const double s1 = slope_of(px[H[nh - 2]], py[H[nh - 2]], px[H[nh - 1]], py[H[nh - 1]]);
                const double s2 = slope_of(px[H[nh - 1]], py[H[nh - 1]], qx, qy);
                if (orient_of(s1, s2) == -1) break;
                nh--;                                                                // pop
                if (px[H[0]] == px[H[nh - 1]] && py[H[0]] == py[H[nh - 1]]) break;  // list equality :58-59
            }
            H[nh++] = (uint8_t)p;
        }
        if (pass == 0) m = nh - 1;  // drop the last of each chain (:89-92)
        else {
#pragma unroll 1
            for (int i = 0; i < nh - 1; i++) hl[m++] = hu[i];
        }
    }
    return m;
}

// scale_down (D:space.py:341-345) + point_in_polygen (convex_hull.py:97-112) fused: the shrunk polygon is
// never stored, its vertices are produced on the fly from the hull indices.
__device__ __noinline__ bool pip_shrunk(const double *px, const double *py, const uint8_t *hull, int m, double lat, double lon) {
    double sx = 0, sy = 0;
#pragma unroll 1
    for (int i = 0; i < m; i++) { sx += px[hull[i]]; sy += py[hull[i]]; }
    const double cx = ddiv(sx, (double)m), cy = ddiv(sy, (double)m);
    double jx, jy;
    {
        double v = px[hull[m - 1]], d = v - cx;
        jx = v - d * 0.1;
        v = py[hull[m - 1]]; d = v - cy;
        jy = v - d * 0.1;
    }
    bool odd = false;
#pragma unroll 1
    for (int i = 0; i < m; i++) {
        double v = px[hull[i]], d = v - cx;
        const double ix = v - d * 0.1;
        v = py[hull[i]]; d = v - cy;
        const double iy = v - d * 0.1;
        const double a0 = ix - lat, a1 = iy - lon;
        const double b0 = lat - jx, b1 = lon - jy;
        const double m1 = a0 * b1, m2 = a1 * b0;
        if (m1 - m2 == 0) return false;
        if ((iy < lon && jy >= lon) || (jy < lon && iy >= lon)) {
            const double t = ddiv(lon - iy, jy - iy);
            const double u = t * (jx - ix);
            if (ix + u < lat) odd = !odd;
        }
        jx = ix; jy = iy;
    }
    return odd;
}

// Minimum-norm least squares (np.linalg.lstsq restatement: streaming Givens QR + one-sided Jacobi SVD,
// identical operation order to oracle/pct_oracle_common.h po_ls_*).  ld = leading dimension of R and V.
struct LsWork { double *R, *V, *y, *row, *x; int ld; };

__device__ __noinline__ void ls_add_row(const LsWork &w, int k, double rhs) {
#pragma unroll 1
    for (int i = 0; i < k; i++) {
        const double b = w.row[i];
        if (b == 0) continue;
        const double a = w.R[i * w.ld + i];
        const double r = dsqrt(a * a + b * b);
        const double c = ddiv(a, r), sn = ddiv(b, r);
#pragma unroll 1
        for (int j = i; j < k; j++) {
            const double rij = w.R[i * w.ld + j], vj = w.row[j];
            w.R[i * w.ld + j] = c * rij + sn * vj;
            w.row[j] = c * vj - sn * rij;
        }
        const double yi = w.y[i];
        w.y[i] = c * yi + sn * rhs;
        rhs = c * rhs - sn * yi;
    }
}
__device__ __noinline__ void ls_solve(const LsWork &w, int k, int rows) {
    double *G = w.R, *V = w.V;
    const int ld = w.ld;
#pragma unroll 1
    for (int i = 0; i < k; i++)
#pragma unroll 1
        for (int j = 0; j < k; j++) V[i * ld + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
#pragma unroll 1
        for (int p = 0; p < k - 1; p++)
#pragma unroll 1
            for (int q = p + 1; q < k; q++) {
                double alpha = 0, beta = 0, gamma = 0;
#pragma unroll 1
                for (int i = 0; i < k; i++) {
                    const double gp = G[i * ld + p], gq = G[i * ld + q];
                    alpha += gp * gp;
                    beta += gq * gq;
                    gamma += gp * gq;
                }
                if (gamma == 0 || fabs(gamma) <= 1e-15 * dsqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = ddiv(beta - alpha, 2.0 * gamma);
                const double t = ddiv(zeta >= 0 ? 1.0 : -1.0, fabs(zeta) + dsqrt(1.0 + zeta * zeta));
                const double c = ddiv(1.0, dsqrt(1.0 + t * t)), sn = c * t;
#pragma unroll 1
                for (int i = 0; i < k; i++) {
                    const double gp = G[i * ld + p], gq = G[i * ld + q];
                    G[i * ld + p] = c * gp - sn * gq;
                    G[i * ld + q] = sn * gp + c * gq;
                    const double vp = V[i * ld + p], vq = V[i * ld + q];
                    V[i * ld + p] = c * vp - sn * vq;
                    V[i * ld + q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double smax = 0;  // singular values (kept in w.row)
#pragma unroll 1
    for (int j = 0; j < k; j++) {
        double a = 0;
#pragma unroll 1
        for (int i = 0; i < k; i++) a += G[i * ld + j] * G[i * ld + j];
        const double s = dsqrt(a);
        w.row[j] = s;
        if (s > smax) smax = s;
    }
    const int M = rows > k ? rows : k;
    const double cutoff = 2.220446049250313e-16 * (double)M * smax;
#pragma unroll 1
    for (int i = 0; i < k; i++) w.x[i] = 0;
#pragma unroll 1
    for (int j = 0; j < k; j++) {
        const double sg = w.row[j];
        if (!(sg > cutoff)) continue;
        double uy = 0;
#pragma unroll 1
        for (int i = 0; i < k; i++) uy += G[i * ld + j] * w.y[i];
        const double coef = ddiv(uy, sg * sg);
#pragma unroll 1
        for (int i = 0; i < k; i++) w.x[i] += V[i * ld + j] * coef;
    }
}

// lstsq load split for k >= 3 supports without a direct edge (D:space.py:134-152 / 231-250):
// c2x/c2y = contact-rectangle centres, (cx,cy) = stack COM; writes the k assignment ratios to w.x
__device__ __noinline__ void lstsq_ratios(const LsWork &w, int k, const double *c2x, const double *c2y, double cx, double cy) {
#pragma unroll 1
    for (int i = 0; i < k; i++) {
        w.y[i] = 0;
#pragma unroll 1
        for (int j = 0; j < k; j++) w.R[i * w.ld + j] = 0;
    }
    int rows = 0;
#pragma unroll 1
    for (int a = 0; a < k - 1; a++)
#pragma unroll 1
        for (int b = a + 1; b < k; b++) {
#pragma unroll 1
            for (int t = 0; t < k; t++) w.row[t] = 0;
            const double lx = c2x[a] - c2x[b], ly = c2y[a] - c2y[b];
            const double molecular = dot2(cx - c2x[a], cy - c2y[a], lx, ly);
            if (molecular != 0) {
                const double r = ddiv(fabs(dot2(cx - c2x[b], cy - c2y[b], lx, ly)), molecular);
                w.row[a] = 1;
                w.row[b] = -r;
            }
            ls_add_row(w, k, 0.0);
            rows++;
        }
#pragma unroll 1
    for (int t = 0; t < k; t++) w.row[t] = 1;
    ls_add_row(w, k, 1.0);
    rows++;
    ls_solve(w, k, rows);
}

// ---------------------------------------------------------------------------------------------
// Geometry policy G (per lane) provides:
//   typedef Node            footprint corner + dims of a box (ints for discrete, doubles for continuous) + mass
//   int  n_boxes() const
//   void node_box(int id, Node &out) const                       (real box id)
//   void centre(const Node&, double &cx,&cy,&cz) const            (D:space.py:35)
//   bool support(const Node &node, int t, double r[4]) const      (is box t a support? contact rect x1,y1,x2,y2)
//   bool strictly_inside(cx, cy, const double r[4]) const         (direct-edge test)
// ---------------------------------------------------------------------------------------------
struct StabFrame {
    Stack4 st;        // the node's stack for this visit
    uint8_t node;     // real box index, or NODE_NEW
    uint8_t base, k, i;
    uint8_t whole;    // 1: whole stack to the single support; 2: direct edge (others zero); 0: (c2d_i, st.cz) split
};
constexpr int NODE_NEW = 255;

// Single support (58 % of all visits): the "hull" of the 4 corners of one contact rectangle.
// With P0=(x1,y1) P1=(x1,y2) P2=(x2,y1) P3=(x2,y2) perturbed by x += y*1e-6 (convex_hull.py:43) and
// P0.x < P1.x < P2.x < P3.x (checked by the caller), ConvexHull's chain scans are fully determined:
//   lower: [P0,P1] -> P2: slope(P0,P1) > 0 > slope(P1,P2) => pop P1 (then first==last => break), push P2;
//          P3: slope(P0,P2) = 0 < slope(P2,P3) => keep                     => [P0,P2,P3] -> drop last
//   upper: [P3,P2] -> P1: slope(P3,P2) > 0 > slope(P2,P1) => pop P2, push P1; P0: slope(P3,P1) = -0 < slope(P1,P0)
//                                                                           => [P3,P1,P0] -> drop last
// so the polygon is [P0,P2,P3,P1]; scale_down + point_in_polygen run on registers, same operation order as
// hull_indices + pip_shrunk.
__device__ __noinline__ bool pip_rect(double x1, double y1, double x2, double y2, double t1, double t2, double lat, double lon) {
    // t1 = y1*1e-6, t2 = y2*1e-6 ; hull order P0,P2,P3,P1
    const double hx[4] = {x1 + t1, x2 + t1, x2 + t2, x1 + t2};
    const double hy[4] = {y1, y1, y2, y2};
    const double sx = ((hx[0] + hx[1]) + hx[2]) + hx[3], sy = ((hy[0] + hy[1]) + hy[2]) + hy[3];
    const double cx = ddiv(sx, 4.0), cy = ddiv(sy, 4.0);
    double vx[4], vy[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double d = hx[i] - cx;
        vx[i] = hx[i] - d * 0.1;
        d = hy[i] - cy;
        vy[i] = hy[i] - d * 0.1;
    }
    bool odd = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int j = (i + 3) & 3;
        const double a0 = vx[i] - lat, a1 = vy[i] - lon;
        const double b0 = lat - vx[j], b1 = lon - vy[j];
        const double m1 = a0 * b1, m2 = a1 * b0;
        if (m1 - m2 == 0) return false;
        if ((vy[i] < lon && vy[j] >= lon) || (vy[j] < lon && vy[i] >= lon)) {
            const double t = ddiv(lon - vy[i], vy[j] - vy[i]);
            const double u = t * (vx[j] - vx[i]);
            if (vx[i] + u < lat) odd = !odd;
        }
    }
    return odd;
}

// The DFS keeps the CURRENT node in registers; frames are pushed to the lane-local stack only for nodes with
// >= 2 supports, and descending into the last (or only) support is a tail call (nothing is left to do in the
// parent once its last child returns True).
template <bool REAL, class G>
__device__ __noinline__ int stability_check(const G &g, const typename G::Node &root, EdgePool &pool, BigScratch *big, int *lock,
                                            const int new_id, int &flags) {
    typedef typename G::Node Node;
    constexpr bool real = REAL;  // REAL: load-propagating update of a committed placement; else read-only feasibility check
    StabFrame fr[STAB_DEPTH];
    uint8_t sup_id[STAB_SUP_POOL];
    double sup_m[STAB_SUP_POOL];
    const int root_id = real ? new_id : NODE_NEW;
    int depth = 0;            // number of pushed frames
    int node = root_id, base = 0;
    Stack4 st;
    g.centre(root, st.cx, st.cy, st.cz);
    st.m = root.mass;

#pragma unroll 1
    for (;;) {
        // ================= ENTER(node, st) =================
        Node cur;
        if (node != root_id) g.node_box(node, cur);
        else cur = root;
        const int limit = (node == root_id) ? g.n_boxes() : node;
        int k = 0, sid0 = 0;
        double r0[4], r[4];
#pragma unroll 1
        for (int t = 0; t < limit; t++) {
            if (!g.support(cur, t, r)) continue;
            if (base + k >= STAB_SUP_POOL || k >= KSUP_MAX) { flags |= PCT_FLAG_SUPPORT_OVERFLOW; return 0; }
            if (k == 0) { sid0 = t; r0[0] = r[0]; r0[1] = r[1]; r0[2] = r[2]; r0[3] = r[3]; }
            sup_id[base + k] = (uint8_t)t;
            k++;
        }
        int child = -1;          // >= 0: tail-descend into this support with load (vx,vy,st.cz,vm)
        double vx = st.cx, vy = st.cy, vm = st.m;
        if (k == 1) {
            const double t1 = r0[1] * 1e-6, t2 = r0[3] * 1e-6;
            const bool fast = (r0[0] + t1 < r0[0] + t2) && (r0[0] + t2 < r0[2] + t1) && (r0[2] + t1 < r0[2] + t2);
            bool ok;
            if (fast) ok = pip_rect(r0[0], r0[1], r0[2], r0[3], t1, t2, st.cx, st.cy);
            else {
                double px[4] = {r0[0] + t1, r0[0] + t2, r0[2] + t1, r0[2] + t2}, py[4] = {r0[1], r0[3], r0[1], r0[3]};
                uint8_t order[4], hl[8], hu[8];
                const int m = hull_indices(px, py, 4, order, hl, hu);
                ok = pip_shrunk(px, py, hl, m, st.cx, st.cy);
            }
            if (!ok) return 0;
            if (real) {
                int pos = -1;
#pragma unroll 1
                for (int q = 0; q < pool.n; q++)
                    if (pool.upper[q] == node && pool.lower[q] == sid0) { pos = q; break; }
                if (pos < 0) {
                    if (pool.n >= EDGE_MAX) { flags |= PCT_FLAG_EDGE_OVERFLOW; return 0; }
                    pos = pool.n++;
                    pool.upper[pos] = (uint8_t)node;
                    pool.lower[pos] = (uint8_t)sid0;
                }
                pool.st[pos] = st;
            }
            child = sid0;  // whole stack goes to the single support
        } else if (k >= 2) {
            // ---------- general case: hull over all contact-rectangle corners ----------
            double lrect[KSUP_SMALL][4], lpx[4 * KSUP_SMALL], lpy[4 * KSUP_SMALL];
            uint8_t lorder[4 * KSUP_SMALL], lhl[8 * KSUP_SMALL], lhu[4 * KSUP_SMALL + 4];
            const bool small = k <= KSUP_SMALL;
            double (*rect)[4] = lrect;
            double *px = lpx, *py = lpy;
            uint8_t *order = lorder, *hl = lhl, *hu = lhu;
            if (!small) {  // rare: serialise the lanes of this env on the per-env HBM scratch
                while (atomicCAS(lock, 0, 1) != 0) { }
                __threadfence_block();
                rect = big->rect; px = big->px; py = big->py; order = big->order; hl = big->hl; hu = big->hu;
            }
#pragma unroll 1
            for (int s = 0; s < k; s++) g.support(cur, sup_id[base + s], rect[s]);
            // combine_contact_points order: (x1,y1) (x1,y2) (x2,y1) (x2,y2); perturb x += y*1e-6 (convex_hull.py:43)
#pragma unroll 1
            for (int s = 0; s < k; s++) {
                const double x1 = rect[s][0], y1 = rect[s][1], x2 = rect[s][2], y2 = rect[s][3];
                const double t1 = y1 * 1e-6, t2 = y2 * 1e-6;
                px[4 * s + 0] = x1 + t1; py[4 * s + 0] = y1;
                px[4 * s + 1] = x1 + t2; py[4 * s + 1] = y2;
                px[4 * s + 2] = x2 + t1; py[4 * s + 2] = y1;
                px[4 * s + 3] = x2 + t2; py[4 * s + 3] = y2;
            }
            const int m = hull_indices(px, py, 4 * k, order, hl, hu);
            bool ok = pip_shrunk(px, py, hl, m, st.cx, st.cy);
            int whole = 2;
            if (ok) {
                // --- distribution ---
                int direct = -1;
#pragma unroll 1
                for (int s = 0; s < k; s++)
                    if (g.strictly_inside(st.cx, st.cy, rect[s])) { direct = s; break; }
                if (direct >= 0) {
#pragma unroll 1
                    for (int s = 0; s < k; s++) sup_m[base + s] = (s == direct) ? st.m : 0.0;
                } else {
                    whole = 0;
                    // contact-rectangle centres (centre2D, D:space.py:371) reuse px/py (hull no longer needed)
#pragma unroll 1
                    for (int s = 0; s < k; s++) {
                        px[s] = (rect[s][0] + rect[s][2]) * 0.5;  // (x1 + x2) / 2, exact
                        py[s] = (rect[s][1] + rect[s][3]) * 0.5;
                    }
                    if (k == 2) {
                        double lx = px[0] - px[1], ly = py[0] - py[1];
                        const double len = dsqrt(fma(ly, ly, lx * lx));
                        const double len2 = len * len;
                        lx = ddiv(lx, len2); ly = ddiv(ly, len2);
                        sup_m[base + 0] = st.m * fabs(dot2(st.cx - px[1], st.cy - py[1], lx, ly));
                        sup_m[base + 1] = st.m * fabs(dot2(st.cx - px[0], st.cy - py[0], lx, ly));
                    } else {
                        double lR[KSUP_SMALL * KSUP_SMALL], lV[KSUP_SMALL * KSUP_SMALL], ly_[KSUP_SMALL], lrow[KSUP_SMALL], lx_[KSUP_SMALL];
                        LsWork w;
                        w.R = small ? lR : big->R; w.V = small ? lV : big->V; w.y = small ? ly_ : big->y;
                        w.row = small ? lrow : big->row; w.x = small ? lx_ : big->x; w.ld = small ? KSUP_SMALL : KSUP_MAX;
                        lstsq_ratios(w, k, px, py, st.cx, st.cy);
#pragma unroll 1
                        for (int s = 0; s < k; s++) sup_m[base + s] = st.m * w.x[s];
                    }
                }
                if (real) {
                    // persist the loads: up_edges[self] = Stack(...) for every support, in support order
#pragma unroll 1
                    for (int s = 0; s < k; s++) {
                        Stack4 e = st;
                        if (!whole) { e.cx = px[s]; e.cy = py[s]; }
                        e.m = sup_m[base + s];
                        int pos = -1;
#pragma unroll 1
                        for (int q = 0; q < pool.n; q++)
                            if (pool.upper[q] == node && pool.lower[q] == sup_id[base + s]) { pos = q; break; }
                        if (pos < 0) {
                            if (pool.n >= EDGE_MAX) { flags |= PCT_FLAG_EDGE_OVERFLOW; ok = false; break; }
                            pos = pool.n++;
                            pool.upper[pos] = (uint8_t)node;
                            pool.lower[pos] = sup_id[base + s];
                        }
                        pool.st[pos] = e;
                    }
                }
            }
            if (!small) { __threadfence_block(); atomicExch(lock, 0); }
            if (!ok) return 0;
            if (depth >= STAB_DEPTH) { flags |= PCT_FLAG_SUPPORT_OVERFLOW; return 0; }
            StabFrame &f = fr[depth++];
            f.st = st; f.node = (uint8_t)node; f.base = (uint8_t)base; f.k = (uint8_t)k; f.i = 0; f.whole = (uint8_t)whole;
        }
        // ================= pick the next node to enter =================
        int parent = node;
        if (child < 0) {
            // k == 0 (return True) or a frame was just pushed: continue with the top frame's next child
            for (;;) {
                if (depth == 0) return 1;
                StabFrame &f = fr[depth - 1];
                if (f.i == f.k) { depth--; continue; }  // all supports passed -> True
                const int s = f.i++;
                child = sup_id[f.base + s];
                parent = f.node;
                st = f.st;
                vm = sup_m[f.base + s];
                vx = st.cx; vy = st.cy;
                if (!f.whole) {  // load sits at the centre of this support's contact rectangle with the parent
                    Node par;
                    if (parent != root_id) g.node_box(parent, par);
                    else par = root;
                    g.support(par, child, r);
                    vx = (r[0] + r[2]) * 0.5; vy = (r[1] + r[3]) * 0.5;
                }
                base = f.base + f.k;
                if (s == f.k - 1) { depth--; base = f.base; }  // tail call: the parent frame is finished
                break;
            }
        }
        // ---- calculate_new_com of `child` (D:space.py:51-71) under the load (vx, vy, st.cz, vm) of `parent` ----
        Node sb;
        g.node_box(child, sb);
        double ccx, ccy, ccz, mm = sb.mass;
        g.centre(sb, ccx, ccy, ccz);
        ccx *= mm; ccy *= mm; ccz *= mm;
#pragma unroll 1
        for (int q = 0; q < pool.n; q++) {
            if (pool.lower[q] != child) continue;
            if (!real && pool.upper[q] == parent) continue;  // `involved` path member: its real load is replaced by the virtual one
            const Stack4 e = pool.st[q];
            ccx += e.cx * e.m; ccy += e.cy * e.m; ccz += e.cz * e.m;
            mm += e.m;
        }
        if (!real && vm != 0.0) {  // zero-mass virtual loads add +0.0 to every sum: skipped (exact)
            ccx += vx * vm; ccy += vy * vm; ccz += st.cz * vm;
            mm += vm;
        }
        st.cx = ddiv(ccx, mm); st.cy = ddiv(ccy, mm); st.cz = ddiv(ccz, mm); st.m = mm;
        node = child;
    }
}

}  // namespace pct
