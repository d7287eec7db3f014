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
//     -fmad=false); fma() appears only where it reproduces OpenBLAS ddot on 2-vectors.
#pragma once
#include <cstdint>
#include <math.h>
#include <math_constants.h>
#include "pct_kernels.h"

namespace pct {

__device__ __forceinline__ double dot2(double u0, double u1, double v0, double v1) { return fma(u1, v1, u0 * v0); }

__device__ __forceinline__ double slope_of(double ax, double ay, double bx, double by) {
    if (bx != ax) return (by - ay) / (bx - ax);
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
    for (int i = 0; i < n; i++) order[i] = (uint8_t)i;
    for (int i = 1; i < n; i++) {  // stable insertion sort by x (sorted(key=x[0]), :34-37)
        uint8_t o = order[i];
        double kx = px[o];
        int j = i - 1;
        while (j >= 0 && px[order[j]] > kx) { order[j + 1] = order[j]; j--; }
        order[j + 1] = o;
    }
    int cnt[2];
    for (int pass = 0; pass < 2; pass++) {
        uint8_t *H = pass == 0 ? hl : hu;
        int nh = 0;
        double s1 = 0, s2 = 0;
        for (int k = 0; k < n; k++) {
            int p = order[pass == 0 ? k : n - 1 - k];
            double qx = px[p], qy = py[p];
            if (nh >= 2) {
                s1 = slope_of(px[H[nh - 2]], py[H[nh - 2]], px[H[nh - 1]], py[H[nh - 1]]);
                s2 = slope_of(px[H[nh - 1]], py[H[nh - 1]], qx, qy);
            }
            while (nh >= 2 && orient_of(s1, s2) != -1) {
                nh--;
                if (px[H[0]] == px[H[nh - 1]] && py[H[0]] == py[H[nh - 1]]) break;  // list equality :58-59
                s1 = slope_of(px[H[nh - 2]], py[H[nh - 2]], px[H[nh - 1]], py[H[nh - 1]]);
                s2 = slope_of(px[H[nh - 1]], py[H[nh - 1]], qx, qy);
            }
            H[nh++] = (uint8_t)p;
        }
        cnt[pass] = nh;
    }
    int m = cnt[0] - 1;
    for (int i = 0; i < cnt[1] - 1; i++) hl[m++] = hu[i];
    return m;
}

// scale_down (D:space.py:341-345) + point_in_polygen (convex_hull.py:97-112) fused: the shrunk polygon is
// never stored, its vertices are produced on the fly from the hull indices.
__device__ __noinline__ bool pip_shrunk(const double *px, const double *py, const uint8_t *hull, int m, double lat, double lon) {
    double sx = 0, sy = 0;
    for (int i = 0; i < m; i++) { sx += px[hull[i]]; sy += py[hull[i]]; }
    double cx = sx / (double)m, cy = sy / (double)m;
    auto vx = [&](int i) { double v = px[hull[i]]; double d = v - cx; return v - d * 0.1; };
    auto vy = [&](int i) { double v = py[hull[i]]; double d = v - cy; return v - d * 0.1; };
    double jx = vx(m - 1), jy = vy(m - 1);
    bool odd = false;
    for (int i = 0; i < m; i++) {
        double ix = vx(i), iy = vy(i);
        double a0 = ix - lat, a1 = iy - lon;
        double b0 = lat - jx, b1 = lon - jy;
        double m1 = a0 * b1, m2 = a1 * b0;
        if (m1 - m2 == 0) return false;
        if ((iy < lon && jy >= lon) || (jy < lon && iy >= lon)) {
            double t = (lon - iy) / (jy - iy);
            double u = t * (jx - ix);
            if (ix + u < lat) odd = !odd;
        }
        jx = ix; jy = iy;
    }
    return odd;
}

// Minimum-norm least squares (np.linalg.lstsq restatement: streaming Givens QR + one-sided Jacobi SVD,
// identical operation order to oracle/pct_oracle_common.h po_ls_*).  ld = leading dimension of R and V.
struct LsWork { double *R, *V, *y, *row, *x; int ld; };

__device__ __noinline__ void ls_init(const LsWork &w, int k) {
    for (int i = 0; i < k; i++) {
        w.y[i] = 0;
        for (int j = 0; j < k; j++) w.R[i * w.ld + j] = 0;
    }
}
__device__ __noinline__ void ls_add_row(const LsWork &w, int k, double rhs) {
    for (int i = 0; i < k; i++) {
        double b = w.row[i];
        if (b == 0) continue;
        double a = w.R[i * w.ld + i];
        double r = sqrt(a * a + b * b);
        double c = a / r, sn = b / r;
        for (int j = i; j < k; j++) {
            double rij = w.R[i * w.ld + j], vj = w.row[j];
            w.R[i * w.ld + j] = c * rij + sn * vj;
            w.row[j] = c * vj - sn * rij;
        }
        double yi = w.y[i];
        w.y[i] = c * yi + sn * rhs;
        rhs = c * rhs - sn * yi;
    }
}
__device__ __noinline__ void ls_solve(const LsWork &w, int k, int rows) {
    double *G = w.R, *V = w.V;
    const int ld = w.ld;
    for (int i = 0; i < k; i++)
        for (int j = 0; j < k; j++) V[i * ld + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (int p = 0; p < k - 1; p++)
            for (int q = p + 1; q < k; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < k; i++) {
                    double gp = G[i * ld + p], gq = G[i * ld + q];
                    alpha += gp * gp;
                    beta += gq * gq;
                    gamma += gp * gq;
                }
                if (gamma == 0 || fabs(gamma) <= 1e-15 * sqrt(alpha * beta)) continue;
                rotated = true;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < k; i++) {
                    double gp = G[i * ld + p], gq = G[i * ld + q];
                    G[i * ld + p] = c * gp - sn * gq;
                    G[i * ld + q] = sn * gp + c * gq;
                    double vp = V[i * ld + p], vq = V[i * ld + q];
                    V[i * ld + p] = c * vp - sn * vq;
                    V[i * ld + q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    // singular values; reuse w.row for sigma
    double smax = 0;
    for (int j = 0; j < k; j++) {
        double a = 0;
        for (int i = 0; i < k; i++) a += G[i * ld + j] * G[i * ld + j];
        double s = sqrt(a);
        w.row[j] = s;
        if (s > smax) smax = s;
    }
    int M = rows > k ? rows : k;
    double cutoff = 2.220446049250313e-16 * (double)M * smax;
    for (int i = 0; i < k; i++) w.x[i] = 0;
    for (int j = 0; j < k; j++) {
        double sg = w.row[j];
        if (!(sg > cutoff)) continue;
        double uy = 0;
        for (int i = 0; i < k; i++) uy += G[i * ld + j] * w.y[i];
        double coef = uy / (sg * sg);
        for (int i = 0; i < k; i++) w.x[i] += V[i * ld + j] * coef;
    }
}

// ---------------------------------------------------------------------------------------------
// Geometry policy G must provide (all per lane):
//   int  n_boxes() const
//   void node_box(int id, StabNode &out) const                                          (real box id)
//   bool support_rect(const StabNode &node, int t, double &x1,&y1,&x2,&y2) const        (is box t a support?)
//   bool strictly_inside(cx, cy, x1,y1,x2,y2) const                                     (direct-edge test)
// ---------------------------------------------------------------------------------------------
struct StabNode { double lx, ly, lz, dx, dy, dz, mass; };  // corner + dims (centre = l + d/2, D:space.py:35)

struct StabFrame {
    Stack4 st;        // this node's (virtual) stack for this visit
    uint8_t node;     // real box index, or NODE_NEW
    uint8_t base, k, i;
    uint8_t whole;    // 1: children receive the whole stack centre (k==1 / direct edge); 0: (c2d_i, st.cz)
};
constexpr int NODE_NEW = 255;

enum StabResult { STAB_FALSE = 0, STAB_TRUE = 1 };

template <bool REAL, class G>
__device__ __noinline__ int stability_check(const G &g, const StabNode &root, EdgePool &pool, BigScratch *big, int *lock,
                                            int new_id, int &flags) {
    // new_id: index the new box gets if REAL placement succeeds (== n_boxes); virtual candidates use NODE_NEW
    StabFrame fr[STAB_DEPTH];
    uint8_t sup_id[STAB_SUP_POOL];
    double sup_m[STAB_SUP_POOL];
    int depth = 0;
    fr[0].st = Stack4{root.lx + root.dx / 2, root.ly + root.dy / 2, root.lz + root.dz / 2, root.mass};
    fr[0].node = REAL ? (uint8_t)new_id : (uint8_t)NODE_NEW;
    fr[0].base = 0; fr[0].k = 0xFF; fr[0].i = 0; fr[0].whole = 1;
    StabNode cur = root;

    while (depth >= 0) {
        StabFrame &f = fr[depth];
        if (f.k == 0xFF) {
            // ---------------- ENTER: supports, hull, PIP, load distribution ----------------
            if (f.node != (REAL ? new_id : NODE_NEW)) g.node_box(f.node, cur);
            else cur = root;
            const int limit = (f.node == (REAL ? new_id : NODE_NEW)) ? g.n_boxes() : (int)f.node;
            int k = 0;
            const int base = f.base;
            bool overflow = false;
            for (int t = 0; t < limit; t++) {
                double x1, y1, x2, y2;
                if (g.support_rect(cur, t, x1, y1, x2, y2)) {
                    if (base + k >= STAB_SUP_POOL || k >= KSUP_MAX) { overflow = true; break; }
                    sup_id[base + k] = (uint8_t)t;
                    k++;
                }
            }
            if (overflow) { flags |= 32; return STAB_FALSE; }
            f.k = (uint8_t)k;
            f.i = 0;
            if (k == 0) { depth--; continue; }  // return True
            // --- polygon test ---
            bool inside;
            {
                double lpx[4 * KSUP_SMALL], lpy[4 * KSUP_SMALL];
                uint8_t lorder[4 * KSUP_SMALL], lhl[8 * KSUP_SMALL], lhu[4 * KSUP_SMALL + 4];
                const bool small = k <= KSUP_SMALL;
                if (!small) { while (atomicCAS(lock, 0, 1) != 0) { } __threadfence_block(); }
                double *px = small ? lpx : big->px, *py = small ? lpy : big->py;
                uint8_t *order = small ? lorder : big->order, *hl = small ? lhl : big->hl, *hu = small ? lhu : big->hu;
                for (int s = 0; s < k; s++) {
                    double x1, y1, x2, y2;
                    g.support_rect(cur, sup_id[base + s], x1, y1, x2, y2);
                    // combine_contact_points order: (x1,y1) (x1,y2) (x2,y1) (x2,y2); perturb x += y*1e-6 (convex_hull.py:43)
                    double t1 = y1 * 1e-6, t2 = y2 * 1e-6;
                    px[4 * s + 0] = x1 + t1; py[4 * s + 0] = y1;
                    px[4 * s + 1] = x1 + t2; py[4 * s + 1] = y2;
                    px[4 * s + 2] = x2 + t1; py[4 * s + 2] = y1;
                    px[4 * s + 3] = x2 + t2; py[4 * s + 3] = y2;
                }
                int m = hull_indices(px, py, 4 * k, order, hl, hu);
                inside = pip_shrunk(px, py, hl, m, f.st.cx, f.st.cy);
                if (!small) { __threadfence_block(); atomicExch(lock, 0); }
            }
            if (!inside) return STAB_FALSE;
            // --- distribution ---
            f.whole = 1;
            if (k == 1) {
                sup_m[base] = f.st.m;
            } else {
                int direct = -1;
                for (int s = 0; s < k; s++) {
                    double x1, y1, x2, y2;
                    g.support_rect(cur, sup_id[base + s], x1, y1, x2, y2);
                    if (g.strictly_inside(f.st.cx, f.st.cy, x1, y1, x2, y2)) { direct = s; break; }
                }
                if (direct >= 0) {
                    for (int s = 0; s < k; s++) sup_m[base + s] = (s == direct) ? f.st.m : 0.0;
                    f.whole = 2;  // non-direct supports get a zero-mass load (centre irrelevant numerically)
                } else if (k == 2) {
                    f.whole = 0;
                    double x1, y1, x2, y2;
                    g.support_rect(cur, sup_id[base + 0], x1, y1, x2, y2);
                    double c0x = (x1 + x2) / 2, c0y = (y1 + y2) / 2;
                    g.support_rect(cur, sup_id[base + 1], x1, y1, x2, y2);
                    double c1x = (x1 + x2) / 2, c1y = (y1 + y2) / 2;
                    double lx = c0x - c1x, ly = c0y - c1y;
                    double len = sqrt(fma(ly, ly, lx * lx));
                    double len2 = len * len;
                    lx = lx / len2; ly = ly / len2;
                    double r0 = fabs(dot2(f.st.cx - c1x, f.st.cy - c1y, lx, ly));
                    double r1 = fabs(dot2(f.st.cx - c0x, f.st.cy - c0y, lx, ly));
                    sup_m[base + 0] = f.st.m * r0;
                    sup_m[base + 1] = f.st.m * r1;
                } else {
                    f.whole = 0;
                    double lR[KSUP_SMALL * KSUP_SMALL], lV[KSUP_SMALL * KSUP_SMALL], ly_[KSUP_SMALL], lrow[KSUP_SMALL], lx_[KSUP_SMALL];
                    const bool small = k <= KSUP_SMALL;
                    if (!small) { while (atomicCAS(lock, 0, 1) != 0) { } __threadfence_block(); }
                    LsWork w;
                    w.R = small ? lR : big->R; w.V = small ? lV : big->V; w.y = small ? ly_ : big->y;
                    w.row = small ? lrow : big->row; w.x = small ? lx_ : big->x; w.ld = small ? KSUP_SMALL : KSUP_MAX;
                    ls_init(w, k);
                    int rows = 0;
                    for (int a = 0; a < k - 1; a++) {
                        double x1, y1, x2, y2;
                        g.support_rect(cur, sup_id[base + a], x1, y1, x2, y2);
                        double cax = (x1 + x2) / 2, cay = (y1 + y2) / 2;
                        for (int b = a + 1; b < k; b++) {
                            g.support_rect(cur, sup_id[base + b], x1, y1, x2, y2);
                            double cbx = (x1 + x2) / 2, cby = (y1 + y2) / 2;
                            for (int t = 0; t < k; t++) w.row[t] = 0;
                            double lx = cax - cbx, ly = cay - cby;
                            double molecular = dot2(f.st.cx - cax, f.st.cy - cay, lx, ly);
                            if (molecular != 0) {
                                double r = fabs(dot2(f.st.cx - cbx, f.st.cy - cby, lx, ly)) / molecular;
                                w.row[a] = 1;
                                w.row[b] = -r;
                            }
                            ls_add_row(w, k, 0.0);
                            rows++;
                        }
                    }
                    for (int t = 0; t < k; t++) w.row[t] = 1;
                    ls_add_row(w, k, 1.0);
                    rows++;
                    ls_solve(w, k, rows);
                    for (int s = 0; s < k; s++) sup_m[base + s] = f.st.m * w.x[s];
                    if (!small) { __threadfence_block(); atomicExch(lock, 0); }
                }
            }
            if (REAL) {
                // persist the loads: up_edges[self] = Stack(...) for every support, in support order
                for (int s = 0; s < k; s++) {
                    Stack4 e;
                    if (f.whole) { e.cx = f.st.cx; e.cy = f.st.cy; e.cz = f.st.cz; }
                    else {
                        double x1, y1, x2, y2;
                        g.support_rect(cur, sup_id[base + s], x1, y1, x2, y2);
                        e.cx = (x1 + x2) / 2; e.cy = (y1 + y2) / 2; e.cz = f.st.cz;
                    }
                    e.m = sup_m[base + s];
                    int pos = -1;
                    for (int q = 0; q < pool.n; q++)
                        if (pool.upper[q] == f.node && pool.lower[q] == sup_id[base + s]) { pos = q; break; }
                    if (pos < 0) {
                        if (pool.n >= EDGE_MAX) { flags |= 16; return STAB_FALSE; }
                        pos = pool.n++;
                        pool.upper[pos] = f.node;
                        pool.lower[pos] = sup_id[base + s];
                    }
                    pool.st[pos] = e;
                }
            }
            continue;  // fall into CHILD phase on the next iteration
        }
        // ---------------- CHILD: recurse into support f.i ----------------
        if (f.i == f.k) { depth--; continue; }  // all supports passed -> True
        if (depth + 1 >= STAB_DEPTH) { flags |= 32; return STAB_FALSE; }
        const int s = f.i++;
        const int sid = sup_id[f.base + s];
        StabNode sb;
        g.node_box(sid, sb);
        // calculate_new_com (D:space.py:51-71)
        double ccx = (sb.lx + sb.dx / 2) * sb.mass, ccy = (sb.ly + sb.dy / 2) * sb.mass, ccz = (sb.lz + sb.dz / 2) * sb.mass;
        double mm = sb.mass;
        for (int q = 0; q < pool.n; q++) {
            if (pool.lower[q] != sid) continue;
            if (!REAL && pool.upper[q] == f.node) continue;  // `involved` path member: its real load is replaced by the virtual one
            Stack4 e = pool.st[q];
            ccx += e.cx * e.m; ccy += e.cy * e.m; ccz += e.cz * e.m;
            mm += e.m;
        }
        if (!REAL) {
            double vm = sup_m[f.base + s];
            if (f.whole != 2 || vm != 0.0) {  // zero-mass loads add +0.0: skipped (exact)
                double vx, vy;
                if (f.whole) { vx = f.st.cx; vy = f.st.cy; }
                else {
                    // centre2D of this support's contact rectangle with the parent
                    StabNode par;
                    if (f.node != NODE_NEW) g.node_box(f.node, par);
                    else par = root;
                    double x1, y1, x2, y2;
                    g.support_rect(par, sid, x1, y1, x2, y2);
                    vx = (x1 + x2) / 2; vy = (y1 + y2) / 2;
                }
                ccx += vx * vm; ccy += vy * vm; ccz += f.st.cz * vm;
                mm += vm;
            }
        }
        ccx /= mm; ccy /= mm; ccz /= mm;
        StabFrame &c = fr[depth + 1];
        c.st = Stack4{ccx, ccy, ccz, mm};
        c.node = (uint8_t)sid;
        c.base = (uint8_t)(f.base + f.k);
        c.k = 0xFF; c.i = 0; c.whole = 1;
        depth++;
    }
    return STAB_TRUE;
}

}  // namespace pct
