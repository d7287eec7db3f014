// Discrete-domain geometry policy of the stability routine (pct_stability.cuh) and the integer resting-height loop: shared by the
// kernels of pct_discrete.cu and by the host build of the routine that the CPU tests drive (tests/host_emul/).
#pragma once
#include <cstdint>
#include "pct_kernels.h"

namespace pct {

struct NodeD { int lx, ly, lz, dx, dy, dz; double mass; };

struct GeomD {
    typedef NodeD Node;
    const int16_t (*box)[6];
    int n;
    const double *den;  // per-box density (setting 3) or nullptr (density 1)
    __device__ __forceinline__ int n_boxes() const { return n; }
    __device__ __forceinline__ void node_box(int id, NodeD &o) const {
        const int16_t *b = box[id];
        o.lx = b[0]; o.ly = b[1]; o.lz = b[2];
        o.dx = b[3] - b[0]; o.dy = b[4] - b[1]; o.dz = b[5] - b[2];
        o.mass = (double)(o.dx * o.dy * o.dz) * (den ? den[id] : 1.0);
    }
    __device__ __forceinline__ void centre(const NodeD &o, double &cx, double &cy, double &cz) const {  // D:space.py:35
        cx = (double)o.lx + (double)o.dx * 0.5;
        cy = (double)o.ly + (double)o.dy * 0.5;
        cz = (double)o.lz + (double)o.dz * 0.5;
    }
    // D:space.py:360-372 — box t supports the node iff its top equals the node's bottom and the footprints
    // overlap with positive area
    __device__ __forceinline__ bool support(const NodeD &nd, int t, double r[4]) const {
        const int16_t *b = box[t];
        if ((int)b[5] != nd.lz) return false;
        const int x1 = max(nd.lx, (int)b[0]), y1 = max(nd.ly, (int)b[1]);
        const int x2 = min(nd.lx + nd.dx, (int)b[3]), y2 = min(nd.ly + nd.dy, (int)b[4]);
        if (x1 >= x2 || y1 >= y2) return false;
        r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2;
        return true;
    }
    __device__ __forceinline__ bool strictly_inside(double cx, double cy, const double r[4]) const {
        return cx > r[0] && cx < r[2] && cy > r[1] && cy < r[3];  // D:space.py:89-90,186-187
    }
};

// resting height of a footprint: max top over the placed boxes it overlaps (== np.max(plain[lx:lx+x, ly:ly+y]))
__device__ __forceinline__ int rest_height(const int16_t (*box)[6], int first, int n, int stride, int lx, int ly, int hx, int hy) {
    int mh = 0;
#pragma unroll 4
    for (int t = first; t < n; t += stride) {
        const int16_t *b = box[t];
        if (lx < b[3] && hx > b[0] && ly < b[4] && hy > b[1]) mh = max(mh, (int)b[5]);
    }
    return mh;
}

// The same loop, also collecting the supports of the footprint (the placed boxes whose top IS the resting height and that overlap it =
// GeomD::support, D:space.py:360-372) in scan order: k = their number, pack = the first four ids (8 bits each).  One pass over the boxes
// instead of the resting-height pass followed by the support scan of the stability routine.
// far_out: the footprint's centre lies outside the bounding box of the contact rectangles by at least half a cell.  The support polygon of
// the placement (convex hull of the contact-rectangle corners, x perturbed by y * 1e-6, shrunk towards its vertex mean; D:space.py:341-345,
// 378-379, convex_hull.py:39-112) lies inside that box up to 2.6e-4, and point_in_polygen answers False for a point outside the polygon's
// bounding box (no edge straddles the ray, or an even number of crossings, or the collinear early-out — all False), so the placement fails
// its root test without any hull being built.  Integer test on doubled coordinates; the margin (0.5) is 2000x the perturbation.
__device__ __forceinline__ int rest_height_supports(const int16_t (*box)[6], int n, int lx, int ly, int hx, int hy, int &k, uint32_t &pack, bool &far_out) {
    int mh = 0, kk = 0;
    uint32_t pk = 0;
    int X1 = 0, Y1 = 0, X2 = 0, Y2 = 0;
#pragma unroll 2
    for (int t = 0; t < n; t++) {
        const int16_t *b = box[t];
        if (lx < b[3] && hx > b[0] && ly < b[4] && hy > b[1]) {
            const int top = b[5];
            if (top > mh) { mh = top; kk = 0; pk = 0; }
            if (top == mh) {
                const int x1 = max(lx, (int)b[0]), y1 = max(ly, (int)b[1]), x2 = min(hx, (int)b[3]), y2 = min(hy, (int)b[4]);
                if (kk == 0) { X1 = x1; Y1 = y1; X2 = x2; Y2 = y2; }
                else { X1 = min(X1, x1); Y1 = min(Y1, y1); X2 = max(X2, x2); Y2 = max(Y2, y2); }
                if (kk < 4) pk |= (uint32_t)t << (8 * kk);
                kk++;
            }
        }
    }
    k = kk; pack = pk;
    const int cx2 = lx + hx, cy2 = ly + hy;  // 2 x centre
    far_out = kk > 0 && (cx2 < 2 * X1 || cx2 > 2 * X2 || cy2 < 2 * Y1 || cy2 > 2 * Y2);
    return mh;
}

}  // namespace pct
