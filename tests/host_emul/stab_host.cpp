// TEST INFRASTRUCTURE — a HOST build of the device stability routine (csrc/pct_stability.cuh + pct_geom.cuh, compiled by g++ with the
// few CUDA intrinsics it uses shimmed) so that the `-m "not gpu"` tests can drive the very source the kernels compile against the CPU
// oracle / the reference's records: every real placement (stability_check<true>: load persistence in the edge pool) and every virtual
// feasibility check (stability_check<false>) of whole trajectories.  It mirrors what pct_apply_kernel does around the call
// (csrc/pct_discrete.cu: CSR slot of the new box, resting height, commit of the box and of the pool counters) and nothing else.
// What it does NOT cover: warp-level code, TMA staging, SASS — those are the GPU tests' business.  g++ -ffp-contract=off == nvcc -fmad=false.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
using std::max;
using std::min;
static inline float __fdividef(float a, float b) { return a / b; }  // device: approximate; only a pre-filter whose verdict the exact path confirms
static inline int atomicCAS(int *p, int cmp, int val) { int o = *p; if (o == cmp) *p = val; return o; }
static inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }
static inline void __threadfence_block() {}
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
#include <cuda_runtime.h>
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#include "pct_stability.cuh"
#include "pct_geom.cuh"

using namespace pct;

struct StabHost {
    int setting, W, L, H;
    int n_box, n_edge, n_poly, flags;
    int16_t box[NB_MAX][6];
    double density[NB_MAX];
    uint8_t e_lower[EDGE_MAX + 1], e_next[EDGE_MAX + 1], first_in[NB_MAX], last_in[NB_MAX];
    uint16_t e_off[NB_MAX + 2], poly_off[NB_MAX + 2];
    Stack4 e_st[EDGE_MAX + 1];
    double poly[POLY_MAX][2];
    BigScratch big;
    int lock;
    // ALIAS variant (EdgePoolA): the reference's object semantics of the load entries
    int alias;
    Stack4 box_st[NB_MAX + 1];
    uint8_t e_upper[EDGE_MAX + 1];
    uint32_t e_alias[(EDGE_MAX + 32) / 32];
};

static EdgePool pool_of(StabHost *h) {
    return EdgePool{h->e_lower, h->e_next, h->e_off, h->first_in, h->last_in, h->e_st, h->e_st, h->n_edge, h->poly_off, &h->poly[0][0], &h->poly[0][0], h->n_poly};
}

extern "C" {
StabHost *sh_create(int setting, int W, int L, int H) {
    StabHost *h = new StabHost();
    memset(h, 0, sizeof *h);
    h->setting = setting; h->W = W; h->L = L; h->H = H;
    return h;
}
void sh_destroy(StabHost *h) { delete h; }
void sh_set_alias(StabHost *h, int on) { h->alias = on; }
void sh_reset(StabHost *h) { h->n_box = 0; h->n_edge = 0; h->n_poly = 0; h->flags = 0; h->lock = 0; }
int sh_flags(StabHost *h) { return h->flags; }
int sh_n_boxes(StabHost *h) { return h->n_box; }

// Space.drop_box_virtual (D:space.py:393-433): feasibility of the oriented item (x, y, z) at (lx, ly); *mh_out = resting height
int sh_virtual(StabHost *h, int x, int y, int z, int lx, int ly, double density, int *mh_out) {
    const int mh = rest_height(h->box, 0, h->n_box, 1, lx, ly, lx + x, ly + y);
    if (mh_out) *mh_out = mh;
    if (lx + x > h->W || ly + y > h->L) return 0;
    if (mh + z > h->H) return 0;
    if (h->setting == 2 || mh == 0) return 1;
    GeomD g{h->box, h->n_box, h->setting == 3 ? h->density : nullptr};
    NodeD root{lx, ly, mh, x, y, z, (double)(x * y * z) * density};
    EdgePool pool = pool_of(h);
    int fl = 0;
    const int ok = stability_check<false, GeomD>(g, root, pool, &h->big, &h->lock, 0, fl) != 0;
    h->flags |= fl;
    return ok;
}

// Space.drop_box (D:space.py:347-391) as pct_apply_kernel performs it: 1 = placed, 0 = rejected (the episode ends)
int sh_place(StabHost *h, int x, int y, int z, int lx, int ly, double density) {
    const int n0 = h->n_box;
    if (n0 >= NB_MAX) return 0;
    h->e_off[n0] = (uint16_t)h->n_edge;
    h->poly_off[n0] = (uint16_t)h->n_poly;
    h->first_in[n0] = EDGE_NIL;
    const int mh = rest_height(h->box, 0, n0, 1, lx, ly, lx + x, ly + y);
    if (lx + x > h->W || ly + y > h->L) return 0;
    if (mh + z > h->H) return 0;
    if (h->setting != 2 && mh != 0) {
        GeomD g{h->box, n0, h->setting == 3 ? h->density : nullptr};
        NodeD root{lx, ly, mh, x, y, z, (double)(x * y * z) * density};
        EdgePoolA pool;
        static_cast<EdgePool &>(pool) = pool_of(h);
        pool.box_st = h->box_st; pool.e_upper = h->e_upper; pool.e_alias = h->e_alias;
        int fl = 0;
        const int res = h->alias ? stability_check<true, GeomD, true>(g, root, pool, &h->big, &h->lock, n0, fl)
                                 : stability_check<true, GeomD>(g, root, pool, &h->big, &h->lock, n0, fl);
        h->n_edge = pool.n;
        h->n_poly = pool.n_poly;
        h->flags |= fl;
        if (!res) return 0;
    }
    int16_t *b = h->box[n0];
    b[0] = (int16_t)lx; b[1] = (int16_t)ly; b[2] = (int16_t)mh; b[3] = (int16_t)(lx + x); b[4] = (int16_t)(ly + y); b[5] = (int16_t)(mh + z);
    h->density[n0] = density;
    h->n_box = n0 + 1;
    h->e_off[n0 + 1] = (uint16_t)h->n_edge;
    h->poly_off[n0 + 1] = (uint16_t)h->n_poly;
    return 1;
}
}
