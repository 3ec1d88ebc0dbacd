// sm_100a device kernels for the cholespy solve path.  Compiled to a cubin, embedded in the host
// library and loaded through the driver API (cuda_driver.cpp).
//
// What they replace (reference: kernels/kernels.cu):
//   chb_analysis            <- analysis_lower / analysis_upper   (:8-33, :35-60)
//   chb_sweep_*  (sweep.cuh)       <- solve_lower_* / solve_upper_*  (:64-131, :133-202)   [default; _exact: reference order]
//   chb_front_*  (front_sweep.cuh) <- the same over the supernodal front image (CHB_SWEEP=front)
//
// chb_analysis: persistent co-resident grid, one warp per row, rows handed out in order by an atomic
// ticket; the level word doubles as the ready flag (pre-set to -1), so a waiting row can never read a
// stale level (the reference publishes level and flag with two unordered stores, which races on a
// B200: profiles/r01_reference_analysis_race.txt).
#include <cstdint>

#include "kernel_args.h"

using chb::AnalysisArgs;

namespace {

constexpr int kChunk = 4;        // rows handed out per ticket
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Hands out the next chunk of positions in the level-ordered row list to the calling warp.
__device__ __forceinline__ int next_chunk(int *ticket, int lane) {
    int t = 0;
    if (lane == 0) t = atomicAdd(ticket, kChunk);
    return __shfl_sync(kFull, t, 0);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Level-set analysis.  level[row] = 0 if the row has no off-diagonal entry, else 1 + max level of
// its dependencies.  `level` must be pre-set to -1; the level word itself is the ready flag.
// descending == 0: CSR of L  (dependencies = all but the LAST entry, rows visited 0..n-1)
// descending == 1: CSR of L' (dependencies = all but the FIRST entry, rows visited n-1..0)
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256) chb_analysis(const AnalysisArgs a) {
    const int lane = threadIdx.x & 31;
    int warp_max = 0;
    for (;;) {
        const int t0 = next_chunk(a.ticket, lane);
        if (t0 >= a.n) break;
        const int t1 = min(t0 + kChunk, a.n);
        for (int t = t0; t < t1; ++t) {
            const int row = a.descending ? a.n - 1 - t : t;
            int beg = a.rowptr[row], end = a.rowptr[row + 1];
            if (a.descending) beg += 1; else end -= 1;
            int lvl = 0;
            for (int e = beg + lane; e < end; e += 32) {
                const int *dep = a.level + a.colidx[e];
                int d;
                unsigned backoff = 32;
                while ((d = ld_relaxed(dep)) < 0) {  // back off: thousands of polling warps would saturate L2
                    __nanosleep(backoff);
                    if (backoff < 1024) backoff *= 2;
                }
                lvl = max(lvl, d + 1);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) lvl = max(lvl, __shfl_xor_sync(kFull, lvl, o));
            if (lane == 0) st_relaxed(a.level + row, lvl);
            warp_max = max(warp_max, lvl);
        }
    }
    if (lane == 0 && warp_max > 0) atomicMax(a.max_level, warp_max);
}

// ------------------------------------------------------------------------------------------------
// Level-set analysis through the ELIMINATION TREE (default).  For a Cholesky factor the level of row i in the forward
// solve (longest dependency path, kernels/kernels.cu:8-33 of the reference) is the HEIGHT of node i in the
// elimination tree, and its level in the backward solve (:35-60) is its DEPTH: row i depends exactly on its
// tree descendants (forward) / ancestors (backward) that appear in its row / column, its children / parent among
// them.  Both are computed for all nodes at once by pointer doubling: after round r every node knows its 2^r-th
// ancestor, the number of edges up to it, and the longest downward path of fewer than 2^r edges below it; one round
// is one launch of n threads, ceil(log2(tree height)) + 1 rounds in all (14 for the 7,008 levels of cfg4) instead of
// a serial chain of thousands of dependent polls.  Bit-exact with chb_analysis (tests/test_gpu_parity.py).
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256) chb_etree_round(const chb::EtreeArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int up = a.anc_in[i];
    int dist = a.dist_in[i];
    const int best = a.best_in[i];
    atomicMax(a.best_out + i, best);
    int up2 = -1;
    if (up >= 0) {
        atomicMax(a.best_out + up, best + a.step);  // the node is exactly `step` edges below `up`
        dist += a.dist_in[up];
        up2 = a.anc_in[up];
    }
    a.anc_out[i] = up2;
    a.dist_out[i] = dist;
    if (up2 >= 0) atomicAdd(a.active, 1);
}

#include "sweep.cuh"
#include "local_sweep.cuh"
#include "wide_sweep.cuh"
#include "front_sweep.cuh"
