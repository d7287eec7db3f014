// The piece queue of the fork-join walk kernels (pct_walk_fork_kernel / pctc_walk_fork_kernel): device-side protocol shared by both domains.
//
// A step's stability walks that the light-prefix kernel could not finish are cut into PIECES (stab_piece, pct_stability.cuh): one chain of visits
// each; a node with k >= 2 supports keeps its first support subtree and publishes the other k - 1 as new pieces.  The queue is one array of
// WalkPiece entries per launch:
//     [0, n_init)            the pieces the light-prefix kernel wrote (it has completed before the consumer kernel starts: plain data, dealt
//                            statically — warp w takes entries [w * L, (w + 1) * L), no atomics)
//     [n_init, n_init + ..)  pieces forked while the kernel runs: producers take slots with an atomicAdd on ctr[ALLOC], write the entry and then set
//                            ready[slot]; consumers take TICKETS with one atomicAdd per warp on ctr[HEAD].  A ticket beyond what has been allocated is a
//                            claim on a FUTURE piece: its lane polls ready[slot] (an address of its own: no hot spot) and is served the moment a walk
//                            forks.  No compare-and-swap loops anywhere (a first version with a CAS pop ran 10-40x slower: retry storms).
// Termination: ctr[OUTSTANDING] counts pieces queued or running (a piece that forks increments it BEFORE it finishes itself, so it can only
// reach zero when no piece can appear any more); whoever brings it to zero sets ctr[DONE], the only word waiting lanes poll besides their slot.
// Only the first `keep` warps hold tickets on future pieces (helpers for the tail); the others leave as soon as the fork queue has no unclaimed entry,
// which frees their SM slots for the emit kernel's blocks.  The last warp out resets every counter for the next step.
#pragma once
#include <cstdint>
#include "pct_kernels.h"
#include "pct_stability.cuh"

namespace pct {

enum { PQ_NINIT = 0, PQ_HEAD = 1, PQ_OUTSTANDING = 2, PQ_EXITED = 3, PQ_ALLOC = 4, PQ_DONE = 5, PQ_WORDS = 8 };

struct PieceQueue {
    WalkPiece *q;
    int32_t *ready;   // [cap] 1 = entry written (set by the producer after a fence, cleared by the consumer)
    int32_t *ctr;     // [PQ_WORDS]
    int cap;
};

// light-prefix kernel: `count` pieces of this warp, returns the first index (the caller writes entry + walk_pend and nothing else: the consumer kernel
// starts after this kernel has completed)
__device__ __forceinline__ int pq_reserve_initial(const PieceQueue &pq, int count) {
    atomicAdd(pq.ctr + PQ_OUTSTANDING, count);
    return atomicAdd(pq.ctr + PQ_NINIT, count);
}

// a running piece publishes a forked one; false = the queue is full (the caller flags the env and fails the walk)
__device__ __forceinline__ bool pq_fork(const PieceQueue &pq, int n_init, int32_t *walk_pend, const WalkPiece &pc) {
    const int slot = n_init + atomicAdd(pq.ctr + PQ_ALLOC, 1);
    if (slot >= pq.cap) return false;
    atomicAdd(walk_pend, 1);                   // before the parent's own decrement: the walk cannot complete in between
    atomicAdd(pq.ctr + PQ_OUTSTANDING, 1);
    pq.q[slot] = pc;
    __threadfence();
    *(volatile int32_t *)(pq.ready + slot) = 1;
    return true;
}

// a piece has finished (after its walk_pend bookkeeping)
__device__ __forceinline__ void pq_piece_done(const PieceQueue &pq) {
    if (atomicSub(pq.ctr + PQ_OUTSTANDING, 1) == 1) { __threadfence(); *(volatile int32_t *)(pq.ctr + PQ_DONE) = 1; }
}

// wait for the forked piece of ticket `t`: true = it is there (entry readable), false = the step's walks are all done
__device__ __forceinline__ bool pq_wait(const PieceQueue &pq, int slot) {
    int spins = 0;
#pragma unroll 1
    for (;;) {
        if (slot < pq.cap && *(volatile const int32_t *)(pq.ready + slot) != 0) { __threadfence(); return true; }
        if ((spins & 7) == 0 && *(volatile const int32_t *)(pq.ctr + PQ_DONE) != 0) {
            // DONE is set after the last piece finished, and a piece is published before its producer finishes: look once more
            if (slot < pq.cap && *(volatile const int32_t *)(pq.ready + slot) != 0) { __threadfence(); return true; }
            return false;
        }
        __nanosleep(spins < 64 ? 100 : 500);
        if (++spins > (1 << 22)) return false;  // bounded (never seen); an orphaned piece shows up as PCT_FLAG_SYNC_TIMEOUT of its env in the emit kernel
    }
}

// last warp out (called by lane 0 of every warp when it leaves)
__device__ __forceinline__ void pq_warp_exit(const PieceQueue &pq, int n_warps, int32_t *walk_ctr) {
    __threadfence();
    if (atomicAdd(pq.ctr + PQ_EXITED, 1) == n_warps - 1) {
        *walk_ctr = 0;
#pragma unroll
        for (int i = 0; i < PQ_WORDS; i++) pq.ctr[i] = 0;
    }
}

}  // namespace pct
