// Discrete PCT environment: batched reset / step kernels for sm_100a.
//
// One warp owns one environment for the whole step:
//   TMA bulk load of the env's packed record (header + placed boxes + EMS list) HBM -> shared memory
//   -> decode action -> real placement (warp max-reduce over placed boxes for the resting height,
//      lane 0 runs the load-propagating stability update) -> EMS update (warp ballot / scan compaction)
//   -> candidate leaves in CPython-set order (hash table in shared memory)
//   -> feasibility, ONE LANE PER CANDIDATE (bounds, resting height over all placed boxes, stacking stability)
//   -> ordered compaction of the first `leaf_node_holder` feasible leaves (ballot + popc)
//   -> coalesced observation write, TMA bulk store of the record back to HBM.
//
// Reference behaviour restated here (D: = pct_envs/PctDiscrete0/ in the reference repo):
//   PackingDiscrete.step / reset / cur_observation / get_possible_position / LeafNode2Action  D:bin3D.py:61-188
//   Space.drop_box / drop_box_virtual / check_box / update_height_graph                       D:space.py:316-454
//   Space.GENEMS / Difference / EliminateInscribedEMS / EMSPoint                               D:space.py:457-570
//   ShmemVecEnv worker auto-reset                                                              wrapper/shmem_vec_env.py:139-143
// The height map of the reference is not materialised: max(plain[lx:lx+x, ly:ly+y]) equals the maximum top
// over placed boxes whose footprint overlaps the query footprint (update_height_graph sets covered cells to
// the new top, which is >= every older value there).
#include "pct_common.cuh"
#include "pct_stability.cuh"
#include "pct_kernels.h"

namespace pct {

// ------------------------------------------------------------------------------------------------------
struct GeomD {
    const int16_t (*box)[6];
    int n;
    const double *den;  // per-box density (setting 3) or nullptr (density 1)
    __device__ __forceinline__ int n_boxes() const { return n; }
    __device__ __forceinline__ void node_box(int id, StabNode &o) const {
        const int16_t *b = box[id];
        int x = b[3] - b[0], y = b[4] - b[1], z = b[5] - b[2];
        o.lx = b[0]; o.ly = b[1]; o.lz = b[2];
        o.dx = x; o.dy = y; o.dz = z;
        o.mass = (double)(x * y * z) * (den ? den[id] : 1.0);
    }
    // D:space.py:360-372 — box t supports the node iff its top equals the node's bottom and the footprints
    // overlap with positive area
    __device__ __forceinline__ bool support_rect(const StabNode &nd, int t, double &x1, double &y1, double &x2, double &y2) const {
        const int16_t *b = box[t];
        if ((double)b[5] != nd.lz) return false;
        x1 = fmax(nd.lx, (double)b[0]);
        y1 = fmax(nd.ly, (double)b[1]);
        x2 = fmin(nd.lx + nd.dx, (double)b[3]);
        y2 = fmin(nd.ly + nd.dy, (double)b[4]);
        return !(x1 >= x2 || y1 >= y2);
    }
    __device__ __forceinline__ bool strictly_inside(double cx, double cy, double x1, double y1, double x2, double y2) const {
        return cx > x1 && cx < x2 && cy > y1 && cy < y2;  // D:space.py:89-90,186-187
    }
};

// ---- shared-memory layout of one warp ------------------------------------------------------------------
constexpr int SM_HOT = 0;
constexpr int SM_X = sizeof(DEnvHot);               // tables / EMS temp
constexpr int SM_TAB_A = SM_X;                      // 2048 x u16
constexpr int SM_TAB_B = SM_X + TAB_A * 2;          // 512 x u16
constexpr int SM_LEAF = SM_TAB_B + TAB_B * 2;       // NL_MAX x 6 x i16
constexpr int SM_MISC = SM_LEAF + NL_MAX * 12;      // mbarrier (8) + lock (4) + pad
constexpr int SM_PER_WARP = SM_MISC + 16;
static_assert(SM_X % 16 == 0 && SM_PER_WARP % 16 == 0, "alignment");
static_assert(E_MAX * 12 <= TAB_A * 2, "EMS temp aliases table A");

constexpr uint16_t EMPTY = 0xFFFF;

struct Item3 { int d[3]; };

// rotation table of D:space.py:540-562
__device__ __forceinline__ bool rot_dims(const int nb[3], int rot, int &sx, int &sy, int &sz) {
    switch (rot) {
    case 0: sx = nb[0]; sy = nb[1]; sz = nb[2]; return true;
    case 1: sx = nb[1]; sy = nb[0]; sz = nb[2]; return sx != sy;
    case 2: sx = nb[0]; sy = nb[2]; sz = nb[1]; return !(sx == sy && sy == sz);
    case 3: sx = nb[1]; sy = nb[2]; sz = nb[0]; return !(sx == sy && sy == sz);
    case 4: sx = nb[2]; sy = nb[0]; sz = nb[1]; return sx != sy;
    default: sx = nb[2]; sy = nb[1]; sz = nb[0]; return sx != sy;
    }
}

// candidate code = ems_idx << 5 | rot << 2 | corner
__device__ __forceinline__ void cand_decode(uint16_t code, const int16_t (*ems)[6], const int nb[3], int &xs, int &ys, int &zs, int &sx,
                                            int &sy, int &sz) {
    const int16_t *m = ems[code >> 5];
    rot_dims(nb, (code >> 2) & 7, sx, sy, sz);
    const int q = code & 3;
    xs = (q & 1) ? m[3] - sx : m[0];
    ys = (q & 2) ? m[4] - sy : m[1];
    zs = m[2];
}
__device__ __forceinline__ uint64_t cand_key(int xs, int ys, int zs, int sx, int sy, int sz) {
    return (uint64_t)xs | ((uint64_t)ys << 10) | ((uint64_t)zs << 20) | ((uint64_t)(xs + sx) << 30) | ((uint64_t)(ys + sy) << 40) |
           ((uint64_t)(zs + sz) << 50);
}
__device__ __forceinline__ uint64_t cand_hash(int xs, int ys, int zs, int sx, int sy, int sz) {
    uint64_t l[6] = {(uint64_t)xs, (uint64_t)ys, (uint64_t)zs, (uint64_t)(xs + sx), (uint64_t)(ys + sy), (uint64_t)(zs + sz)};
    return tuple_hash6(l);
}

// set_insert_clean (setobject.c): executed uniformly by the whole warp (broadcast shared-memory reads)
__device__ __forceinline__ void table_insert_clean(uint16_t *tab, uint32_t mask, uint64_t hash, uint16_t code, int lane) {
    uint64_t perturb = hash;
    uint32_t i = (uint32_t)hash & mask;
    for (;;) {
        const int probes = (i + 9 <= mask) ? 9 : 0;
        for (int j = 0; j <= probes; j++)
            if (tab[i + j] == EMPTY) {
                if (lane == 0) tab[i + j] = code;
                __syncwarp();
                return;
            }
        perturb >>= 5;
        i = (uint32_t)((uint64_t)i * 5 + 1 + perturb) & mask;
    }
}

// ---- EMS update (GENEMS + Difference + EliminateInscribedEMS, D:space.py:457-531) -------------------------
__device__ void genems_warp(int16_t (*ems)[6], int &n_ems, int16_t (*tmp)[6], const int it[6], double low_bound, int lane, int &flags) {
    const int n0 = n_ems;
    const double lb = low_bound == 0 ? 0.1 : low_bound;
    constexpr int CH = E_MAX / 32;
    uint32_t cm[CH];  // bits 0-4 child present, bit 5 intersected (deleted)
    int16_t ch[CH][4];  // x1,x2 / y1,y2 / z2 of the intersection needed by the children: store x1,x2,y1,y2 ; z2 separately
    int16_t chz[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int i = c * 32 + lane;
        cm[c] = 0;
        if (i < n0) {
            const int16_t *m = ems[i];
            int x1 = it[0], y1 = it[1], z1 = it[2], x2 = it[3], y2 = it[4], z2 = it[5];
            if (m[0] > x1) x1 = m[0];
            if (m[1] > y1) y1 = m[1];
            if (m[2] > z1) z1 = m[2];
            if (m[3] < x2) x2 = m[3];
            if (m[4] < y2) y2 = m[4];
            if (m[5] < z2) z2 = m[5];
            if (x1 > x2) x1 = x2;
            if (y1 > y2) y1 = y2;
            if (z1 > z2) z1 = z2;
            if (!(x1 == x2 || y1 == y2 || z1 == z2)) {
                const int a1 = m[0], b1 = m[1], c1 = m[2], a2 = m[3], b2 = m[4], c2 = m[5];
                const bool uy = (double)(b2 - b1) >= lb, uz = (double)(c2 - c1) >= lb, ux = (double)(a2 - a1) >= lb;
                uint32_t k = 32;
                if ((double)(x1 - a1) >= lb && uy && uz) k |= 1;
                if ((double)(a2 - x2) >= lb && uy && uz) k |= 2;
                if (ux && (double)(y1 - b1) >= lb && uz) k |= 4;
                if (ux && (double)(b2 - y2) >= lb && uz) k |= 8;
                if (ux && uy && (double)(c2 - z2) >= lb) k |= 16;
                cm[c] = k;
                ch[c][0] = (int16_t)x1; ch[c][1] = (int16_t)x2; ch[c][2] = (int16_t)y1; ch[c][3] = (int16_t)y2;
                chz[c] = (int16_t)z2;
            }
        }
    }
    // survivors, original order
    int off = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int i = c * 32 + lane;
        const bool keep = i < n0 && !(cm[c] & 32);
        const uint32_t bm = __ballot_sync(FULL, keep);
        if (keep) {
            const int p = off + __popc(bm & ((1u << lane) - 1));
#pragma unroll
            for (int t = 0; t < 6; t++) tmp[p][t] = ems[i][t];
        }
        off += __popc(bm);
    }
    // children, parent order then fixed child order (left, right, front, back, top)
    bool overflow = false;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int i = c * 32 + lane;
        const int cnt = __popc(cm[c] & 31);
        const int incl = warp_incl_scan(cnt, lane);
        int p = off + incl - cnt;
        if (cnt) {
            const int16_t *m = ems[i];
            const int16_t a1 = m[0], b1 = m[1], c1 = m[2], a2 = m[3], b2 = m[4], c2 = m[5];
            const int16_t x1 = ch[c][0], x2 = ch[c][1], y1 = ch[c][2], y2 = ch[c][3], z2 = chz[c];
            auto put = [&](int16_t q0, int16_t q1, int16_t q2, int16_t q3, int16_t q4, int16_t q5) {
                if (p < E_MAX) { tmp[p][0] = q0; tmp[p][1] = q1; tmp[p][2] = q2; tmp[p][3] = q3; tmp[p][4] = q4; tmp[p][5] = q5; }
                else overflow = true;
                p++;
            };
            if (cm[c] & 1) put(a1, b1, c1, x1, b2, c2);
            if (cm[c] & 2) put(x2, b1, c1, a2, b2, c2);
            if (cm[c] & 4) put(a1, b1, c1, a2, y1, c2);
            if (cm[c] & 8) put(a1, y2, c1, a2, b2, c2);
            if (cm[c] & 16) put(a1, b1, z2, a2, b2, c2);
        }
        off += __shfl_sync(FULL, incl, 31);
    }
    if (__any_sync(FULL, overflow)) flags |= PCT_FLAG_EMS_OVERFLOW_;
    int n = off < E_MAX ? off : E_MAX;
    __syncwarp();
    // EliminateInscribedEMS: drop i if some j != i contains it (non-strict)
    int w = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int i = c * 32 + lane;
        bool keep = false;
        int16_t a[6];
        if (i < n) {
#pragma unroll
            for (int t = 0; t < 6; t++) a[t] = tmp[i][t];
            keep = true;
            for (int j = 0; j < n; j++) {
                if (j == i) continue;
                const int16_t *b = tmp[j];
                if (a[0] >= b[0] && a[1] >= b[1] && a[2] >= b[2] && a[3] <= b[3] && a[4] <= b[4] && a[5] <= b[5]) { keep = false; break; }
            }
        }
        const uint32_t bm = __ballot_sync(FULL, keep);
        if (keep) {
            const int p = w + __popc(bm & ((1u << lane) - 1));
#pragma unroll
            for (int t = 0; t < 6; t++) ems[p][t] = a[t];
        }
        w += __popc(bm);
    }
    n_ems = w;
    __syncwarp();
}

// ---- candidate leaves in CPython set order (EMSPoint, D:space.py:534-570) ---------------------------------
// returns the candidate count; the ordered codes end up at the start of the returned buffer
__device__ int build_candidates(const int16_t (*ems)[6], int n_ems, const int nb[3], int R, uint16_t *tabA, uint16_t *tabB, uint16_t *&out,
                                int lane, int &flags) {
    uint16_t *tab = tabA;
    uint32_t mask = 7;
    int fill = 0;
    if (lane < 8) tab[lane] = EMPTY;
    __syncwarp();
    const int raw = n_ems * R * 4;
    bool stop = false;
    for (int base = 0; base < raw && !stop; base += 32) {
        const int r = base + lane;
        bool valid = false;
        uint64_t hash = 0, key = 0;
        uint16_t code = 0;
        if (r < raw) {
            const int q = r & 3, er = r >> 2;
            const int rot = er % R, ei = er / R;
            int sx, sy, sz;
            if (rot_dims(nb, rot, sx, sy, sz)) {
                const int16_t *m = ems[ei];
                if (m[3] - m[0] >= sx && m[4] - m[1] >= sy && m[5] - m[2] >= sz) {
                    valid = true;
                    code = (uint16_t)((ei << 5) | (rot << 2) | q);
                    const int xs = (q & 1) ? m[3] - sx : m[0];
                    const int ys = (q & 2) ? m[4] - sy : m[1];
                    key = cand_key(xs, ys, m[2], sx, sy, sz);
                    hash = cand_hash(xs, ys, m[2], sx, sy, sz);
                }
            }
        }
        uint32_t vm = __ballot_sync(FULL, valid);
        while (vm) {
            const int k = __ffs(vm) - 1;
            vm &= vm - 1;
            const uint64_t h = __shfl_sync(FULL, hash, k);
            const uint64_t ky = __shfl_sync(FULL, key, k);
            const uint16_t cd = (uint16_t)__shfl_sync(FULL, (int)code, k);
            // set_add_entry (setobject.c), uniform across the warp
            uint64_t perturb = h;
            uint32_t i = (uint32_t)h & mask;
            bool inserted = false;
            for (bool done = false; !done;) {
                const int probes = (i + 9 <= mask) ? 9 : 0;
                for (int j = 0; j <= probes; j++) {
                    const uint16_t e = tab[i + j];
                    if (e == EMPTY) {
                        if (lane == 0) tab[i + j] = cd;
                        inserted = true;
                        done = true;
                        break;
                    }
                    int xs, ys, zs, sx, sy, sz;
                    cand_decode(e, ems, nb, xs, ys, zs, sx, sy, sz);
                    if (cand_key(xs, ys, zs, sx, sy, sz) == ky) { done = true; break; }
                }
                if (!done) {
                    perturb >>= 5;
                    i = (uint32_t)((uint64_t)i * 5 + 1 + perturb) & mask;
                }
            }
            __syncwarp();
            if (inserted) {
                fill++;
                if ((uint32_t)fill * 5 >= mask * 3) {
                    // set_table_resize(used * 4): smallest power of two > 4 * used, re-insert in slot order
                    uint32_t newsize = 8;
                    while (newsize <= (uint32_t)fill * 4) newsize <<= 1;
                    if (newsize > TAB_A) { flags |= PCT_FLAG_CAND_OVERFLOW_; stop = true; break; }
                    uint16_t *nt = (tab == tabA) ? tabB : tabA;
                    for (uint32_t t = lane; t < newsize; t += 32) nt[t] = EMPTY;
                    __syncwarp();
                    for (uint32_t b2 = 0; b2 <= mask; b2 += 32) {
                        const uint32_t s = b2 + lane;
                        const uint16_t e = s <= mask ? tab[s] : EMPTY;
                        uint64_t eh = 0;
                        if (e != EMPTY) {
                            int xs, ys, zs, sx, sy, sz;
                            cand_decode(e, ems, nb, xs, ys, zs, sx, sy, sz);
                            eh = cand_hash(xs, ys, zs, sx, sy, sz);
                        }
                        uint32_t em = __ballot_sync(FULL, e != EMPTY);
                        while (em) {
                            const int kk = __ffs(em) - 1;
                            em &= em - 1;
                            table_insert_clean(nt, newsize - 1, __shfl_sync(FULL, eh, kk), (uint16_t)__shfl_sync(FULL, (int)e, kk), lane);
                        }
                    }
                    tab = nt;
                    mask = newsize - 1;
                }
            }
        }
    }
    __syncwarp();
    // iteration order = slot order: compact the codes in place
    int cnt = 0;
    for (uint32_t b2 = 0; b2 <= mask; b2 += 32) {
        const uint32_t s = b2 + lane;
        const uint16_t e = s <= mask ? tab[s] : EMPTY;
        const uint32_t em = __ballot_sync(FULL, e != EMPTY);
        __syncwarp();
        if (e != EMPTY) tab[cnt + __popc(em & ((1u << lane) - 1))] = e;
        cnt += __popc(em);
        __syncwarp();
    }
    out = tab;
    return cnt;
}

// ---- item source ------------------------------------------------------------------------------------------
__device__ __forceinline__ void draw_item(const DParams &p, int e, DHdr &h) {
    const uint64_t gid = (uint64_t)(p.env_id_base + e);
    const uint64_t d = (uint64_t)h.draw_pos;
    if (p.item_mode == 0) {
        const uint64_t idx = rnd_u64(p.seed, gid, d) % (uint64_t)p.n_items;
        h.next_box[0] = (int)p.item_set[idx * 3 + 0];
        h.next_box[1] = (int)p.item_set[idx * 3 + 1];
        h.next_box[2] = (int)p.item_set[idx * 3 + 2];
        h.next_den = p.setting == 3 ? rnd_density(p.seed, gid, d) : 1.0;
    } else {
        const double *it = p.stream + ((size_t)e * p.stream_len + (size_t)(d % (uint64_t)p.stream_len)) * 4;
        h.next_box[0] = (int)it[0];
        h.next_box[1] = (int)it[1];
        h.next_box[2] = (int)it[2];
        h.next_den = p.setting == 3 ? it[3] : 1.0;
    }
    h.draw_pos++;
}

__device__ __forceinline__ void reset_space(DEnvHot *hot, const DParams &p, int lane) {
    if (lane == 0) {
        hot->h.n_box = 0;
        hot->h.n_ems = 1;
        hot->h.n_leaf = 0;
        hot->h.flags = 0;
        hot->h.n_edge = 0;
        hot->h.vol_sum = 0;
        hot->h.ep_len = 0;
        hot->h.ep_reward = 0;
        hot->ems[0][0] = 0; hot->ems[0][1] = 0; hot->ems[0][2] = 0;
        hot->ems[0][3] = (int16_t)p.W; hot->ems[0][4] = (int16_t)p.L; hot->ems[0][5] = (int16_t)p.H;
    }
    __syncwarp();
}

template <typename OT>
__device__ void write_obs(const DParams &p, int e, const DEnvHot *hot, const DEnvCold *cold, const int16_t (*leaf)[6], int n_leaf, int lane) {
    OT *obs = (OT *)p.obs + (size_t)e * (size_t)((p.nb + p.nl + 1) * 9);
    const int n_box = hot->h.n_box;
    const int total = (p.nb + p.nl + 1) * 9;
    int s0 = hot->h.next_box[0], s1 = hot->h.next_box[1], s2 = hot->h.next_box[2];
    if (s1 < s0) { int t = s0; s0 = s1; s1 = t; }
    if (s2 < s1) { int t = s1; s1 = s2; s2 = t; }
    if (s1 < s0) { int t = s0; s0 = s1; s1 = t; }
    for (int f = lane; f < total; f += 32) {
        const int row = f / 9, col = f - row * 9;
        double v = 0;
        if (row < p.nb) {
            if (row < n_box) {
                if (col < 6) v = hot->box[row][col];
                else if (col == 6) v = p.setting == 3 ? cold->density[row] : 1.0;
                else if (col == 8) v = 1;
            } else if (row == 0 && col == 8) v = 1;  // D:space.py:294-295
        } else if (row < p.nb + p.nl) {
            const int k = row - p.nb;
            if (k < n_leaf) {
                if (col < 5) v = leaf[k][col];
                else if (col == 5) v = p.H;  // D:bin3D.py:128 — bin height, not ze
                else if (col == 8) v = 1;
            }
        } else {
            if (col == 0) v = hot->h.next_den;
            else if (col == 3) v = s0;
            else if (col == 4) v = s1;
            else if (col == 5) v = s2;
            else if (col == 8) v = 1;
        }
        obs[f] = (OT)v;
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------
template <typename OT, bool STAB>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) pct_discrete_kernel(const DParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * WARPS_PER_BLOCK + warp;
    if (e >= p.n_envs) return;
    unsigned char *sm = smem_raw + (size_t)warp * SM_PER_WARP;
    DEnvHot *hot = (DEnvHot *)(sm + SM_HOT);
    uint16_t *tabA = (uint16_t *)(sm + SM_TAB_A), *tabB = (uint16_t *)(sm + SM_TAB_B);
    int16_t (*ems_tmp)[6] = (int16_t (*)[6])(sm + SM_TAB_A);
    int16_t (*leaf)[6] = (int16_t (*)[6])(sm + SM_LEAF);
    uint64_t *mbar = (uint64_t *)(sm + SM_MISC);
    int *lock = (int *)(sm + SM_MISC + 8);
    DEnvHot *ghot = p.hot + e;
    DEnvCold *cold = p.cold + e;
    EdgePool pool{cold->e_upper, cold->e_lower, cold->e_st, 0};

    if (lane == 0) *lock = 0;
    float reward = 0.f;
    int done = 0;
    pct_step_info_ info{};

    if (p.mode == 0) {
        // ---------------- reset (D:bin3D.py:61-67, D:space.py:290-314) ----------------
        for (int t = lane; t < (int)(sizeof(DEnvHot) / 4); t += 32) ((uint32_t *)hot)[t] = 0;
        __syncwarp();
        reset_space(hot, p, lane);
        if (lane == 0) draw_item(p, e, hot->h);
        __syncwarp();
    } else {
        // ---------------- stage the env record: HBM -> smem via TMA ----------------
        if (lane == 0) {
            mbar_init(mbar, 1);
            fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(mbar, (uint32_t)sizeof(DEnvHot));
            tma_load_1d(hot, ghot, (uint32_t)sizeof(DEnvHot), mbar);
        }
        mbar_wait(mbar, 0);
        __syncwarp();

        DHdr &h = hot->h;
        int nb3[3] = {h.next_box[0], h.next_box[1], h.next_box[2]};
        const int n_box0 = h.n_box, n_leaf0 = h.n_leaf, flags0 = h.flags;
        const double next_den0 = h.next_den;
        __syncwarp();
        // ---- LeafNode2Action (D:bin3D.py:139-149) ----
        int lx = 0, ly = 0, x = nb3[0], y = nb3[1], z = nb3[2];
        bool bad = false;
        {
            double a[6];
            bool zero = true;
            if (p.leaf_idx) {
                const int k = p.leaf_idx[e];
                if (k >= 0 && k < n_leaf0) {
                    zero = false;
                    for (int t = 0; t < 6; t++) a[t] = cold->leaf[k][t];
                }
            } else {
                double s = 0;
                for (int t = 0; t < 6; t++) {
                    a[t] = p.action_f64 ? ((const double *)p.actions)[(size_t)e * 9 + t] : (double)((const float *)p.actions)[(size_t)e * 9 + t];
                    s += a[t];
                }
                zero = (s == 0);
            }
            if (!zero) {
                x = (int)(a[3] - a[0]);
                y = (int)(a[4] - a[1]);
                int rem[3] = {nb3[0], nb3[1], nb3[2]};
                int n = 3;
                bool found = false;
                for (int t = 0; t < n; t++)
                    if (rem[t] == x) { for (int u = t; u < n - 1; u++) rem[u] = rem[u + 1]; n--; found = true; break; }
                if (!found) bad = true;
                found = false;
                for (int t = 0; t < n; t++)
                    if (rem[t] == y) { for (int u = t; u < n - 1; u++) rem[u] = rem[u + 1]; n--; found = true; break; }
                if (!found) bad = true;
                z = rem[0];
                lx = (int)a[0];
                ly = (int)a[1];
            }
        }
        // ---- Space.drop_box (D:space.py:347-389) ----
        bool ok = !bad && lx >= 0 && ly >= 0 && lx < max(p.W, p.L) && ly < max(p.W, p.L) && x > 0 && y > 0;
        int max_h = 0;
        if (ok) {
            // resting height: warp max-reduce over the placed boxes whose footprint overlaps
            int mh = 0;
            for (int t = lane; t < n_box0; t += 32) {
                const int16_t *b = hot->box[t];
                if (lx < b[3] && lx + x > b[0] && ly < b[4] && ly + y > b[1]) mh = max(mh, (int)b[5]);
            }
            max_h = __reduce_max_sync(FULL, mh);
            if (lx + x > p.W || ly + y > p.L) ok = false;
            else if (max_h + z > p.H) ok = false;
            else if (STAB && max_h != 0) {
                int res = 0, fl = 0;
                if (lane == 0) {
                    GeomD g{hot->box, n_box0, p.setting == 3 ? cold->density : nullptr};
                    StabNode root{(double)lx, (double)ly, (double)max_h, (double)x, (double)y, (double)z, (double)(x * y * z) * next_den0};
                    pool.n = h.n_edge;
                    res = stability_check<true>(g, root, pool, &cold->big, lock, n_box0, fl);
                    h.n_edge = pool.n;
                    h.flags |= fl;
                }
                __syncwarp();
                res = __shfl_sync(FULL, res, 0);
                ok = res != 0;
            }
            if (ok && n_box0 >= p.nb) {
                ok = false;
                if (lane == 0) h.flags |= PCT_FLAG_BOX_OVERFLOW_;
            }
        }
        if (bad && lane == 0) h.flags |= PCT_FLAG_BAD_ACTION_;
        __syncwarp();
        const double binvol = (double)(p.W * p.L * p.H);
        if (ok) {
            const int bi = n_box0;
            if (lane == 0) {
                int16_t *b = hot->box[bi];
                b[0] = (int16_t)lx; b[1] = (int16_t)ly; b[2] = (int16_t)max_h;
                b[3] = (int16_t)(lx + x); b[4] = (int16_t)(ly + y); b[5] = (int16_t)(max_h + z);
                if (p.setting == 3) cold->density[bi] = next_den0;
                h.n_box = bi + 1;
                h.vol_sum += x * y * z;
            }
            __syncwarp();
            const int it[6] = {lx, ly, max_h, lx + x, ly + y, max_h + z};
            int n_ems = h.n_ems, fl = 0;
            __syncwarp();
            genems_warp(hot->ems, n_ems, ems_tmp, it, p.low_bound, lane, fl);
            const double rw = (double)(nb3[0] * nb3[1] * nb3[2]) / binvol * 10;  // D:bin3D.py:180-183
            reward = (float)rw;
            if (lane == 0) {
                h.n_ems = n_ems;
                h.flags |= fl;
                h.ep_len++;
                h.ep_reward += rw;
                draw_item(p, e, h);
            }
            info.counter = bi + 1;
            info.flags = flags0 | fl;
            __syncwarp();
        } else {
            // terminal step (D:bin3D.py:160-165) followed by the worker's auto-reset (shmem_vec_env.py:141-142)
            done = 1;
            info.counter = n_box0;
            info.flags = h.flags | (bad ? PCT_FLAG_BAD_ACTION_ : 0);
            info.ratio = (float)((double)h.vol_sum / binvol);
            info.ep_reward = (float)h.ep_reward;
            info.ep_len = h.ep_len + 1;
            __syncwarp();
            reset_space(hot, p, lane);
            if (lane == 0) draw_item(p, e, hot->h);
            __syncwarp();
        }
    }

    // ---------------- cur_observation (D:bin3D.py:70-93) + get_possible_position (:100-136) ----------------
    {
        DHdr &h = hot->h;
        const int nb3[3] = {h.next_box[0], h.next_box[1], h.next_box[2]};
        const int R = p.setting == 2 ? 6 : 2;
        uint16_t *cand = nullptr;
        int fl = 0;
        const int n_cand = build_candidates(hot->ems, h.n_ems, nb3, R, tabA, tabB, cand, lane, fl);
        int n_leaf = 0;
        GeomD g{hot->box, h.n_box, p.setting == 3 ? cold->density : nullptr};
        pool.n = h.n_edge;
        __syncwarp();
        for (int base = 0; base < n_cand && n_leaf < p.nl; base += 32) {
            const int c = base + lane;
            bool feas = false;
            int xs = 0, ys = 0, zs = 0, sx = 0, sy = 0, sz = 0;
            if (c < n_cand) {
                cand_decode(cand[c], hot->ems, nb3, xs, ys, zs, sx, sy, sz);
                // drop_box_virtual (D:space.py:393-433) + check_box (:436-454)
                int mh = 0;
                for (int t = 0; t < h.n_box; t++) {
                    const int16_t *b = hot->box[t];
                    if (xs < b[3] && xs + sx > b[0] && ys < b[4] && ys + sy > b[1]) mh = max(mh, (int)b[5]);
                }
                if (xs + sx > p.W || ys + sy > p.L || xs < 0 || ys < 0) feas = false;
                else if (mh + sz > p.H) feas = false;
                else if (!STAB || mh == 0) feas = true;
                else {
                    StabNode root{(double)xs, (double)ys, (double)mh, (double)sx, (double)sy, (double)sz, (double)(sx * sy * sz) * h.next_den};
                    feas = stability_check<false>(g, root, pool, &cold->big, lock, 0, fl) != 0;
                }
            }
            const uint32_t fm = __ballot_sync(FULL, feas);
            if (feas) {
                const int k = n_leaf + __popc(fm & ((1u << lane) - 1));
                if (k < p.nl) {
                    leaf[k][0] = (int16_t)xs; leaf[k][1] = (int16_t)ys; leaf[k][2] = (int16_t)zs;
                    leaf[k][3] = (int16_t)(xs + sx); leaf[k][4] = (int16_t)(ys + sy); leaf[k][5] = (int16_t)(zs + sz);
                }
            }
            n_leaf += __popc(fm);
        }
        if (n_leaf > p.nl) n_leaf = p.nl;
        fl = __reduce_or_sync(FULL, fl);
        __syncwarp();
        if (lane == 0) {
            h.n_leaf = n_leaf;
            h.n_cand = n_cand;
            h.flags |= fl;
        }
        info.n_leaf = n_leaf;
        info.n_cand = n_cand;
        info.n_ems = h.n_ems;
        // persist the emitted leaves for the leaf-index action path
        for (int t = lane; t < n_leaf * 6; t += 32) ((int16_t *)cold->leaf)[t] = ((int16_t *)leaf)[t];
        __syncwarp();
        write_obs<OT>(p, e, hot, cold, leaf, n_leaf, lane);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (p.reward) p.reward[e] = reward;
        if (p.done) p.done[e] = (uint8_t)done;
        if (p.info) p.info[e] = info;
        // record back to HBM: smem -> global via TMA bulk store
        tma_store_1d(ghot, hot, (uint32_t)sizeof(DEnvHot));
        tma_store_commit_wait();
    }
}

// uniform-random valid-leaf policy: reads only the record headers
__global__ void pct_policy_random_kernel(const DEnvHot *hot, int n_envs, int64_t env_id_base, uint64_t seed, int64_t t, int32_t *leaf_idx) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    const int n = hot[e].h.n_leaf;
    leaf_idx[e] = n > 0 ? (int32_t)(rnd_u64(seed, (uint64_t)(env_id_base + e), (uint64_t)t) % (uint64_t)n) : 0;
}

// ---- launchers ---------------------------------------------------------------------------------------------
size_t discrete_smem_bytes() { return (size_t)SM_PER_WARP * WARPS_PER_BLOCK; }

template <typename OT, bool STAB>
static cudaError_t launch_t(const DParams &p, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = discrete_smem_bytes();
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(pct_discrete_kernel<OT, STAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        attr_set = true;
    }
    const int blocks = (p.n_envs + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    pct_discrete_kernel<OT, STAB><<<blocks, 32 * WARPS_PER_BLOCK, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_discrete(const DParams &p, cudaStream_t st) {
    const bool stab = p.setting != 2;
    if (p.obs_f64) return stab ? launch_t<double, true>(p, st) : launch_t<double, false>(p, st);
    return stab ? launch_t<float, true>(p, st) : launch_t<float, false>(p, st);
}

cudaError_t launch_policy_random_discrete(const DEnvHot *hot, int n_envs, int64_t env_id_base, uint64_t seed, int64_t t, int32_t *leaf_idx,
                                          cudaStream_t st) {
    pct_policy_random_kernel<<<(n_envs + 127) / 128, 128, 0, st>>>(hot, n_envs, env_id_base, seed, t, leaf_idx);
    return cudaGetLastError();
}

}  // namespace pct
