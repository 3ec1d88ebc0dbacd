// sm_100a device kernels for the cholespy solve path.  Compiled to a cubin, embedded in the host
// library and loaded through the driver API (cuda_driver.cpp).
//
// What they replace (reference: kernels/kernels.cu):
//   chb_analysis            <- analysis_lower / analysis_upper   (:8-33, :35-60)
//   chb_solve_wide_*        <- solve_lower_* / solve_upper_*     (:64-131, :133-202), n_rhs > 8
//   chb_solve_narrow_*      <- same, n_rhs <= 8
//
// Design (row-task generation of the path):
//   * persistent, co-resident grid; one WARP per row; rows are handed out in level order in small
//     chunks by one atomic per chunk (the reference uses one 128-thread CTA per row and one atomic per
//     row);
//   * dependencies are per-row ready flags holding the EPOCH of the sweep that produced them
//     (ld.acquire.gpu / st.release.gpu), so nothing has to be cleared between sweeps (the reference
//     clears n flags + the ticket before each of the two launches, src/cholesky_solver.cpp:381-382);
//   * factor entries are fetched 32 at a time, coalesced, and broadcast with warp shuffles; every
//     lane polls a different dependency flag, so a chunk of 32 dependencies costs one L2 round trip
//     (the reference polls one flag per non-zero from thread 0 between two block barriers);
//   * wide variant: lanes own RHS columns (register blocked, up to 4 per lane -> 128), accumulation
//     is sequential in the reference's column order with the same fused multiply-subtract, so
//     results are bit-identical to the reference kernels; narrow variant: lanes own non-zeros and
//     the row dot-product is a shuffle reduction;
//   * P is fused into the forward load, P' into the backward store, as in the reference.
#include <cstdint>

#include "kernel_args.h"

using chb::AnalysisArgs;
using chb::SolveArgs;

namespace {

constexpr int kChunk = 4;        // rows handed out per ticket
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(int *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float fnma(float a, float b, float c) { return fmaf(-a, b, c); }
__device__ __forceinline__ double fnma(double a, double b, double c) { return fma(-a, b, c); }

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
                while ((d = ld_relaxed(dep)) < 0) {}
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
// Triangular sweeps.
//   lower (upper == 0): tmp[r,:] = (src[perm[r],:] - sum_{c<r} L[r,c] tmp[c,:]) / L[r,r]
//   upper (upper == 1): tmp[r,:] = (tmp[r,:]      - sum_{c>r} L[c,r] tmp[c,:]) / L[r,r];  dst[perm[r],:] = tmp[r,:]
// CSR operand: lower = CSR of L (diag last), upper = CSR of L' (diag first, walked backwards).
// ------------------------------------------------------------------------------------------------
// Wide: lane l owns RHS columns l, l+32, ... (NV of them).
template <typename Float, int NV>
__device__ __forceinline__ void solve_wide(const SolveArgs<Float> &a) {
    const int lane = threadIdx.x & 31;
    const int k = a.k;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.ticket_next = 0;
    bool own[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) own[v] = lane + 32 * v < k;

    for (;;) {
        const int t0 = next_chunk(a.ticket, lane);
        if (t0 >= a.n) break;
        const int t1 = min(t0 + kChunk, a.n);
        for (int t = t0; t < t1; ++t) {
            const int row = a.order[t];
            const int beg = a.rowptr[row], end = a.rowptr[row + 1];
            const int cnt = end - beg - 1;  // off-diagonal entries
            const Float diag = a.vals[a.upper ? beg : end - 1];
            Float acc[NV];
            {
                const Float *s = a.upper ? a.tmp + (size_t) row * k : a.src + (size_t) a.perm[row] * k;
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] = own[v] ? (a.upper ? __ldcg(s + lane + 32 * v) : s[lane + 32 * v]) : Float(0);
            }
            for (int base = 0; base < cnt; base += 32) {
                const int pos = base + lane;
                int my_col = 0;
                Float my_val = Float(0);
                if (pos < cnt) {
                    const int e = a.upper ? end - 1 - pos : beg + pos;
                    my_col = a.colidx[e];
                    my_val = a.vals[e];
                    const int *flag = a.ready + my_col;
                    while (ld_acquire(flag) != a.epoch) {}
                }
                __syncwarp();
                const int m = min(32, cnt - base);
                for (int j = 0; j < m; ++j) {
                    const int c = __shfl_sync(kFull, my_col, j);
                    const Float w = __shfl_sync(kFull, my_val, j);
                    const Float *xr = a.tmp + (size_t) c * k + lane;
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        if (own[v]) acc[v] = fnma(w, __ldcg(xr + 32 * v), acc[v]);
                }
            }
            {
                Float *o = a.tmp + (size_t) row * k;
                Float *u = a.upper ? a.dst + (size_t) a.perm[row] * k : nullptr;
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (own[v]) {
                        const Float r = acc[v] / diag;
                        __stcg(o + lane + 32 * v, r);
                        if (a.upper) u[lane + 32 * v] = r;
                    }
            }
            __syncwarp();
            if (lane == 0) st_release(a.ready + row, a.epoch);
        }
    }
}

// Narrow: k <= KMAX <= 8; lane j owns non-zeros j, j+32, ... of the row and all k columns of them.
template <typename Float, int KMAX>
__device__ __forceinline__ void solve_narrow(const SolveArgs<Float> &a) {
    const int lane = threadIdx.x & 31;
    const int k = a.k;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.ticket_next = 0;

    for (;;) {
        const int t0 = next_chunk(a.ticket, lane);
        if (t0 >= a.n) break;
        const int t1 = min(t0 + kChunk, a.n);
        for (int t = t0; t < t1; ++t) {
            const int row = a.order[t];
            const int beg = a.rowptr[row], end = a.rowptr[row + 1];
            const int cnt = end - beg - 1;
            const Float diag = a.vals[a.upper ? beg : end - 1];
            Float rhs = Float(0);
            if (lane < k) rhs = a.upper ? __ldcg(a.tmp + (size_t) row * k + lane) : a.src[(size_t) a.perm[row] * k + lane];
            Float acc[KMAX];
#pragma unroll
            for (int v = 0; v < KMAX; ++v) acc[v] = Float(0);
            for (int pos = lane; pos < cnt; pos += 32) {
                const int e = a.upper ? end - 1 - pos : beg + pos;
                const int c = a.colidx[e];
                const Float w = a.vals[e];
                const int *flag = a.ready + c;
                while (ld_acquire(flag) != a.epoch) {}
                const Float *xr = a.tmp + (size_t) c * k;
#pragma unroll
                for (int v = 0; v < KMAX; ++v)
                    if (v < k) acc[v] = fnma(w, __ldcg(xr + v), acc[v]);
            }
            Float mine = Float(0);
#pragma unroll
            for (int v = 0; v < KMAX; ++v) {
                Float s = acc[v];
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
                if (lane == v) mine = s;
            }
            if (lane < k) {
                const Float r = (rhs + mine) / diag;
                __stcg(a.tmp + (size_t) row * k + lane, r);
                if (a.upper) a.dst[(size_t) a.perm[row] * k + lane] = r;
            }
            __syncwarp();
            if (lane == 0) st_release(a.ready + row, a.epoch);
        }
    }
}

#define CHB_WIDE(NAME, T, NV) \
    extern "C" __global__ void __launch_bounds__(256) NAME(const SolveArgs<T> a) { solve_wide<T, NV>(a); }
#define CHB_NARROW(NAME, T, KM) \
    extern "C" __global__ void __launch_bounds__(256) NAME(const SolveArgs<T> a) { solve_narrow<T, KM>(a); }

CHB_WIDE(chb_solve_wide_f32_1, float, 1)
CHB_WIDE(chb_solve_wide_f32_2, float, 2)
CHB_WIDE(chb_solve_wide_f32_3, float, 3)
CHB_WIDE(chb_solve_wide_f32_4, float, 4)
CHB_WIDE(chb_solve_wide_f64_1, double, 1)
CHB_WIDE(chb_solve_wide_f64_2, double, 2)
CHB_WIDE(chb_solve_wide_f64_3, double, 3)
CHB_WIDE(chb_solve_wide_f64_4, double, 4)
CHB_NARROW(chb_solve_narrow_f32_1, float, 1)
CHB_NARROW(chb_solve_narrow_f32_2, float, 2)
CHB_NARROW(chb_solve_narrow_f32_4, float, 4)
CHB_NARROW(chb_solve_narrow_f32_8, float, 8)
CHB_NARROW(chb_solve_narrow_f64_1, double, 1)
CHB_NARROW(chb_solve_narrow_f64_2, double, 2)
CHB_NARROW(chb_solve_narrow_f64_4, double, 4)
CHB_NARROW(chb_solve_narrow_f64_8, double, 8)
