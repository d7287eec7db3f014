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

}  // namespace pct
