// Internal: device record layouts and kernel launch parameters shared by the kernels and the C-ABI layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "pct_b200.h"

namespace pct {

constexpr int NB_MAX = 80;    // internal_node_holder capacity
constexpr int NL_MAX = 64;    // leaf_node_holder capacity
constexpr int E_MAX = 128;    // EMS list capacity (reference: unbounded python list; max observed 51)
constexpr int TAB_A = 2048;   // CPython-set emulation: largest table (<= 1228 distinct candidates)
constexpr int TAB_B = 512;
constexpr int WARPS_PER_BLOCK = 2;

constexpr int PCT_FLAG_BOX_OVERFLOW_ = PCT_FLAG_BOX_OVERFLOW;
constexpr int PCT_FLAG_BAD_ACTION_ = PCT_FLAG_BAD_ACTION;
constexpr int PCT_FLAG_EMS_OVERFLOW_ = PCT_FLAG_EMS_OVERFLOW;
constexpr int PCT_FLAG_CAND_OVERFLOW_ = PCT_FLAG_CAND_OVERFLOW;
typedef pct_step_info pct_step_info_;

constexpr int KSUP_SMALL = 8;    // supports handled with lane-local scratch
constexpr int KSUP_MAX = 32;     // supports handled with the per-env scratch in HBM (serialised by a lock)
constexpr int STAB_DEPTH = 14;   // DFS depth (levels of boxes on top of each other)
constexpr int STAB_SUP_POOL = 48;
constexpr int EDGE_MAX = 256;

struct Stack4 { double cx, cy, cz, m; };

struct EdgePool {          // per-env, global memory
    uint8_t *upper;        // [EDGE_MAX]
    uint8_t *lower;        // [EDGE_MAX]
    Stack4 *st;            // [EDGE_MAX]
    int n;                 // current count (lane-local copy; the REAL path writes it back)
};

// per-env HBM scratch for the rare big cases (k > KSUP_SMALL)
struct BigScratch {
    double rect[KSUP_MAX][4];
    double px[4 * KSUP_MAX], py[4 * KSUP_MAX];
    uint8_t order[4 * KSUP_MAX], hl[8 * KSUP_MAX], hu[4 * KSUP_MAX + 4];
    double R[KSUP_MAX * KSUP_MAX], V[KSUP_MAX * KSUP_MAX], y[KSUP_MAX], row[KSUP_MAX], x[KSUP_MAX];
};


// ---- discrete domain -----------------------------------------------------------------------------------
// "Hot" record: everything a step reads and rewrites, one contiguous 16-byte-aligned blob per env so that a
// single TMA bulk copy stages it into shared memory and another one writes it back.
struct alignas(16) DHdr {  // 64 bytes
    int32_t n_box, n_ems, n_leaf, flags;
    int64_t draw_pos;      // draws consumed from the item source (one per reset + one per placed box)
    double ep_reward;      // Monitor: sum of rewards of the running episode
    int32_t next_box[3];
    int32_t n_edge;        // load edges in the pool (stability settings)
    double next_den;
    int32_t vol_sum;       // sum of packed volumes (get_ratio numerator)
    int32_t ep_len;
    int32_t n_cand;
    int32_t pad_;
};
struct alignas(16) DEnvHot {
    DHdr h;
    int16_t box[NB_MAX][6];  // lx,ly,lz,hx,hy,hz  (placement order)
    int16_t ems[E_MAX][6];   // x1,y1,z1,x2,y2,z2  (reference list order)
};
static_assert(sizeof(DHdr) == 72 || sizeof(DHdr) == 80 || sizeof(DHdr) == 64, "header size");
static_assert(sizeof(DEnvHot) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

// "Cold" record: touched only by the paths that need it (leaf-index actions, setting-3 densities, the
// stability load edges, the rare >8-support scratch).  Lives in HBM / L2, never staged.
struct DEnvCold {
    int16_t leaf[NL_MAX][6];  // leaves emitted with the last observation (xs,ys,zs,xe,ye,ze)
    double density[NB_MAX];   // per placed box (setting 3)
    uint8_t e_upper[EDGE_MAX], e_lower[EDGE_MAX];
    Stack4 e_st[EDGE_MAX];
    uint32_t cand[1232];      // ordered candidate keys written by K2, read by K3 (<= 1228 distinct candidates)
    uint32_t tab_big[TAB_A];  // 2048-slot stage of the set emulation when it does not live in shared memory
    BigScratch big;
};

struct DParams {
    DEnvHot *hot;
    DEnvCold *cold;
    int n_envs;
    int W, L, H, nb, nl, setting;
    double low_bound;
    int item_mode;
    const double *item_set;
    int n_items;
    const double *stream;
    int stream_len;
    uint64_t seed;
    int64_t env_id_base;
    const void *actions;
    int action_f64;
    const int32_t *leaf_idx;
    void *obs;
    int obs_f64;
    float *reward;
    uint8_t *done;
    pct_step_info *info;
    int mode;  // 0 = reset all, 1 = step
    long long *dbg;  // phase timers (only with -DPCT_PHASE_TIMERS)
};

int discrete_kernels_per_step();
cudaError_t launch_discrete(const DParams &p, cudaStream_t st);
cudaError_t launch_policy_random_discrete(const DEnvHot *hot, int n_envs, int64_t env_id_base, uint64_t seed, int64_t t, int32_t *leaf_idx,
                                          cudaStream_t st);

}  // namespace pct
