// Panel sweep kernels (sm_100a): the forward (L y = P b) and backward (L' x = y, x -> P' x) triangular
// solves over the panel / block / task image built by panel_schedule.cpp.
//
// What they replace: solve_lower_* / solve_upper_* of the reference (kernels/kernels.cu:64-131,
// 133-202: one 128-thread CTA per ROW, one flag poll + two block barriers per non-zero).
//
// Execution model
//   * persistent, co-resident grid; every WARP is an independent worker with a private shared-memory
//     ring of pages and its own mbarriers;
//   * workers draw tickets from one atomic counter; ticket order is a topological order of the task
//     graph (dependencies first), so a worker only ever waits on tickets that are already running;
//   * a task's slab (record headers + factor values, one contiguous byte range) is streamed into the
//     ring by 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), up to kStages pages
//     ahead of the parser; factor bytes are read from HBM exactly once per sweep;
//   * the tile (<= 16 rows x n_rhs) of the panel being solved lives in registers: lane l owns RHS
//     columns l, l+32, ... (NV per lane), so every gathered row of already-solved unknowns is one
//     coalesced 128-byte line per column group and is reused for up to 16 x NV FMAs;
//   * factor values are broadcast from shared memory (LDS.128 for the dense block layout);
//   * dependencies: one epoch-valued ready flag per TASK (ld.acquire.gpu / st.release.gpu); flags
//     are never cleared (the epoch is the solve counter);
//   * P is fused into the forward tile load, P' into the backward tile store.
#pragma once
#include <cstdint>

#include "kernel_args.h"

namespace sweep {

using chb::SweepArgs;
using chb::SweepTask;

constexpr int W = 16;            // panel width == tile rows
constexpr int kStages = 3;       // pages in flight per warp
constexpr int kWarpsPerCta = 4;
constexpr unsigned kFullMask = 0xffffffffu;

enum : uint32_t { REC_NOP = 0, REC_BEGIN_RHS, REC_BEGIN_ZERO, REC_BEGIN_SLOT, REC_ADD_SLOT, REC_BLOCK_D4, REC_BLOCK_D1,
                  REC_BLOCK_TM, REC_END_FINAL, REC_END_SLOT };

template <typename F> struct PageOf { static constexpr int bytes = sizeof(F) == 8 ? 4096 : 2048; };

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float fnma(float a, float b, float c) { return fmaf(-a, b, c); }
__device__ __forceinline__ double fnma(double a, double b, double c) { return fma(-a, b, c); }

// ---- tile helpers --------------------------------------------------------------------------------
template <typename F, int NV>
struct Tile {
    F a[W][NV];
};

// acc[t] for a warp-uniform, run-time t (registers cannot be indexed dynamically: jump table)
#define SWEEP_ROW_CASES(OP) \
    switch (t) {            \
        case 0: OP(0) break; case 1: OP(1) break; case 2: OP(2) break; case 3: OP(3) break;     \
        case 4: OP(4) break; case 5: OP(5) break; case 6: OP(6) break; case 7: OP(7) break;     \
        case 8: OP(8) break; case 9: OP(9) break; case 10: OP(10) break; case 11: OP(11) break; \
        case 12: OP(12) break; case 13: OP(13) break; case 14: OP(14) break; default: OP(15) break; \
    }

template <typename F, int NV>
__device__ __forceinline__ void get_row(const Tile<F, NV> &T, int t, F (&r)[NV]) {
#define SWEEP_GET(I) { _Pragma("unroll") for (int v = 0; v < NV; ++v) r[v] = T.a[I][v]; }
    SWEEP_ROW_CASES(SWEEP_GET)
#undef SWEEP_GET
}
template <typename F, int NV>
__device__ __forceinline__ void put_row(Tile<F, NV> &T, int t, const F (&r)[NV]) {
#define SWEEP_PUT(I) { _Pragma("unroll") for (int v = 0; v < NV; ++v) T.a[I][v] = r[v]; }
    SWEEP_ROW_CASES(SWEEP_PUT)
#undef SWEEP_PUT
}

template <typename F, int NV>
__device__ __forceinline__ void load_row_cg(F (&r)[NV], const F *row, int lane, int k) {
#pragma unroll
    for (int v = 0; v < NV; ++v) r[v] = lane + 32 * v < k ? __ldcg(row + lane + 32 * v) : F(0);
}

template <typename F> struct GroupOf { static constexpr int rows = sizeof(F) == 8 ? 2 : 4; };  // source rows fetched together

// Calls body(x, rank) once per source row of `mask` in sweep order (ascending rows forward,
// descending backward); rows are fetched kGroup at a time, one group ahead of the arithmetic.
template <typename F, int NV, bool UPPER, typename Body>
__device__ __forceinline__ void for_sources(uint32_t mask, const F *rows, int lane, int k, Body &&body) {
    constexpr int kGroup = GroupOf<F>::rows;
    F xa[kGroup][NV], xb[kGroup][NV];
    auto fetch = [&](F(&x)[kGroup][NV]) {
        int cnt = 0;
#pragma unroll
        for (int g = 0; g < kGroup; ++g)
            if (mask) {
                const int s = UPPER ? 31 - __clz(mask) : __ffs(mask) - 1;
                mask &= ~(1u << s);
                load_row_cg<F, NV>(x[g], rows + (size_t) s * k, lane, k);
                cnt = g + 1;
            }
        return cnt;
    };
    auto apply = [&](const F(&x)[kGroup][NV], int cnt, int rank) {
#pragma unroll
        for (int g = 0; g < kGroup; ++g)
            if (g < cnt) body(x[g], rank + g);
    };
    int rank = 0;
    int na = fetch(xa);
    while (na) {
        const int nb = fetch(xb);
        apply(xa, na, rank);
        rank += na;
        if (!nb) break;
        na = fetch(xa);
        apply(xb, nb, rank);
        rank += nb;
    }
}

template <typename F> struct Vec4;
template <> struct Vec4<float> {
    float x[4];
    __device__ __forceinline__ static Vec4 load(const float *p) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        return Vec4{{q.x, q.y, q.z, q.w}};
    }
};
template <> struct Vec4<double> {
    double x[4];
    __device__ __forceinline__ static Vec4 load(const double *p) {
        const double2 q0 = *reinterpret_cast<const double2 *>(p), q1 = *reinterpret_cast<const double2 *>(p + 2);
        return Vec4{{q0.x, q0.y, q1.x, q1.y}};
    }
};

// ---- the worker ----------------------------------------------------------------------------------
template <typename F, int NV, bool UPPER>
__device__ __forceinline__ void sweep_worker(const SweepArgs<F> &a) {
    constexpr int PAGE = PageOf<F>::bytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *ring = smem_raw + (size_t) warp * kStages * PAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t) kWarpsPerCta * kStages * PAGE) + warp * kStages;
    const int k = a.k;
    const int epoch = a.epoch;

    if (blockIdx.x == 0 && threadIdx.x == 0) *a.ticket_next = 0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    uint32_t pages_issued = 0, pages_consumed = 0;  // running counts over all tasks of this warp (warp-uniform)
    Tile<F, NV> T;
    uint32_t last_ready = 0;  // wait code most recently confirmed

    uint32_t peek_code = 0;   // flag fetched ahead of the block that needs it
    int peek_value = 0;

    auto wait_for = [&](uint32_t code) {
        if (code == 0 || code == last_ready) return;
        if (code == peek_code && peek_value == epoch) {
            last_ready = code;
            return;
        }
        const int *flag = a.done + (code - 1);
        // a.reserved (tuning): 0 = acquire-load spin; 1 = acquire first, relaxed spin, acquire last;
        // 2 = relaxed spin only (every mutable datum is read with an L2-coherent ld.cg behind the
        // control dependency on the flag); +4 = back off with nanosleep while spinning
        const int mode = a.reserved & 3;
        if (mode == 0) {
            while (ld_acquire_gpu(flag) != epoch) {
                if (a.reserved & 4) __nanosleep(40);
            }
        } else if (mode == 1) {
            if (ld_acquire_gpu(flag) != epoch) {
                while (ld_relaxed_gpu(flag) != epoch) {
                    if (a.reserved & 4) __nanosleep(40);
                }
                (void) ld_acquire_gpu(flag);
            }
        } else {
            while (ld_relaxed_gpu(flag) != epoch) {
                if (a.reserved & 4) __nanosleep(40);
            }
        }
        last_ready = code;
    };

    for (;;) {
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(a.ticket, 1);
        ticket = __shfl_sync(kFullMask, ticket, 0);
        if (ticket >= a.n_tasks) break;
        const int task = a.order[ticket];
        const SweepTask td = a.tasks[task];
        const unsigned char *base = a.slab + td.offset;
        const uint32_t total = td.bytes;
        const uint32_t n_pages = (total + PAGE - 1) / PAGE;

        auto issue = [&](uint32_t page) {  // lane 0 only
            const uint32_t bytes = min((uint32_t) PAGE, total - page * PAGE);
            const uint32_t stage = pages_issued % kStages;
            mbar_expect_tx(bars + stage, bytes);
            tma_load_1d(ring + (size_t) stage * PAGE, base + (size_t) page * PAGE, bytes, bars + stage);
        };
        const uint32_t ahead = min(n_pages, (uint32_t) kStages);
        for (uint32_t pg = 0; pg < ahead; ++pg) {
            if (lane == 0) issue(pg);
            pages_issued += 1;
        }

        int row0 = 0, w = 0;
        for (uint32_t pg = 0; pg < n_pages; ++pg) {
            const uint32_t stage = pages_consumed % kStages;
            mbar_wait(bars + stage, (pages_consumed / kStages) & 1);
            const unsigned char *p = ring + (size_t) stage * PAGE;
            const unsigned char *pend = p + min((uint32_t) PAGE, total - pg * PAGE);
            while (p < pend) {
                const uint4 h = *reinterpret_cast<const uint4 *>(p);
                const uint32_t type = h.x & 0xff;
                const uint32_t len = (h.x >> 8) * 16;
                const unsigned char *pay = p + 16;
                p += len;
                if (type == REC_BLOCK_D4 || type == REC_BLOCK_D1 || type == REC_BLOCK_TM) {
                    wait_for(h.w);
                    // start fetching the ready flag of the next block while this one is processed
                    if (p < pend) {
                        const uint4 hn = *reinterpret_cast<const uint4 *>(p);
                        const uint32_t tn = hn.x & 0xff;
                        if ((tn == REC_BLOCK_D4 || tn == REC_BLOCK_D1 || tn == REC_BLOCK_TM) && hn.w != 0 && hn.w != last_ready) {
                            peek_code = hn.w;
                            peek_value = ld_acquire_gpu(a.done + (hn.w - 1));
                        }
                    }
                    const uint32_t tmask = h.z >> 16, smask = h.z & 0xffff;
                    const F *val = reinterpret_cast<const F *>(pay);
                    const F *rows = a.tmp + (size_t) (int) h.y * k;
                    if (type == REC_BLOCK_D4) {
                        const int stride = (32 - __clz(tmask) + 3) & ~3;
                        for_sources<F, NV, UPPER>(smask, rows, lane, k, [&](const F(&x)[NV], int rank) {
                            const F *vr = val + rank * stride;
#pragma unroll
                            for (int t4 = 0; t4 < W / 4; ++t4) {
                                if (t4 * 4 < stride) {
                                    const Vec4<F> vv = Vec4<F>::load(vr + t4 * 4);
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
#pragma unroll
                                        for (int v = 0; v < NV; ++v) T.a[t4 * 4 + j][v] = fnma(vv.x[j], x[v], T.a[t4 * 4 + j][v]);
                                }
                            }
                        });
                    } else if (type == REC_BLOCK_D1) {
                        const int nt = __popc(tmask);
                        for_sources<F, NV, UPPER>(smask, rows, lane, k, [&](const F(&x)[NV], int rank) {
                            const F *vr = val + rank * nt;
#pragma unroll
                            for (int t = 0; t < 3; ++t)
                                if (t < nt) {
                                    const F l = vr[t];
#pragma unroll
                                    for (int v = 0; v < NV; ++v) T.a[t][v] = fnma(l, x[v], T.a[t][v]);
                                }
                        });
                    } else {
                        // target-major: sources are fetched kGroup rows at a time and applied to every
                        // target of the block before the next group (same per-target order as row by row)
                        constexpr int kGroup = GroupOf<F>::rows;
                        const int ns = __popc(smask);
                        uint32_t sm = smask;
                        int s_rank = 0;
                        while (sm) {
                            F xs[kGroup][NV];
                            int cnt = 0;
#pragma unroll
                            for (int g = 0; g < kGroup; ++g)
                                if (sm) {
                                    const int s = UPPER ? 31 - __clz(sm) : __ffs(sm) - 1;
                                    sm &= ~(1u << s);
                                    load_row_cg<F, NV>(xs[g], rows + (size_t) s * k, lane, k);
                                    cnt = g + 1;
                                }
                            uint32_t tm = tmask;
                            int t_rank = 0;
                            while (tm) {
                                const int t = __ffs(tm) - 1;
                                tm &= tm - 1;
                                F r[NV];
                                get_row<F, NV>(T, t, r);
                                const F *vr = val + t_rank * ns + s_rank;
#pragma unroll
                                for (int g = 0; g < kGroup; ++g)
                                    if (g < cnt) {
                                        const F l = vr[g];
#pragma unroll
                                        for (int v = 0; v < NV; ++v) r[v] = fnma(l, xs[g][v], r[v]);
                                    }
                                put_row<F, NV>(T, t, r);
                                ++t_rank;
                            }
                            s_rank += cnt;
                        }
                    }
                } else if (type == REC_BEGIN_RHS) {
                    row0 = (int) h.y;
                    w = (int) h.z;
                    const int *perm = reinterpret_cast<const int *>(pay);
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        const F *row = nullptr;
                        if (i < w) row = UPPER ? a.tmp + (size_t) (row0 + i) * k : a.src + (size_t) perm[i] * k;
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            F val = F(0);
                            if (i < w && lane + 32 * v < k) val = UPPER ? __ldcg(row + lane + 32 * v) : __ldg(row + lane + 32 * v);
                            T.a[i][v] = val;
                        }
                    }
                } else if (type == REC_END_FINAL) {
                    const int *perm = reinterpret_cast<const int *>(pay);
                    const F *tv = reinterpret_cast<const F *>(pay + (UPPER ? 4 * W : 0));
#pragma unroll
                    for (int step = 0; step < W; ++step) {
                        const int s = UPPER ? W - 1 - step : step;
                        if (s < w) {
                            const F d = *tv++;
#pragma unroll
                            for (int v = 0; v < NV; ++v) T.a[s][v] = T.a[s][v] / d;
#pragma unroll
                            for (int t = UPPER ? 0 : s + 1; t < (UPPER ? s : W); ++t) {
                                if (UPPER || t < w) {
                                    const F l = *tv++;
#pragma unroll
                                    for (int v = 0; v < NV; ++v) T.a[t][v] = fnma(l, T.a[s][v], T.a[t][v]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        if (i < w) {
                            F *row = a.tmp + (size_t) (row0 + i) * k;
                            F *out = UPPER ? a.dst + (size_t) perm[i] * k : nullptr;
#pragma unroll
                            for (int v = 0; v < NV; ++v)
                                if (lane + 32 * v < k) {
                                    __stcg(row + lane + 32 * v, T.a[i][v]);
                                    if (UPPER) out[lane + 32 * v] = T.a[i][v];
                                }
                        }
                    }
                } else if (type == REC_BEGIN_ZERO) {
                    row0 = (int) h.y;
                    w = (int) h.z;
#pragma unroll
                    for (int i = 0; i < W; ++i)
#pragma unroll
                        for (int v = 0; v < NV; ++v) T.a[i][v] = F(0);
                } else if (type == REC_BEGIN_SLOT || type == REC_ADD_SLOT) {
                    wait_for(h.w);
                    if (type == REC_BEGIN_SLOT) {
                        row0 = (int) h.y;
                        w = (int) (h.z & 0xff);
                    }
                    const F *slot = a.scratch + (size_t) (h.z >> 8) * W * k;
#pragma unroll
                    for (int i = 0; i < W; ++i)
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const F val = lane + 32 * v < k ? __ldcg(slot + (size_t) i * k + lane + 32 * v) : F(0);
                            T.a[i][v] = type == REC_BEGIN_SLOT ? val : T.a[i][v] + val;
                        }
                } else if (type == REC_END_SLOT) {
                    F *slot = a.scratch + (size_t) (h.z >> 8) * W * k;
#pragma unroll
                    for (int i = 0; i < W; ++i)
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            if (lane + 32 * v < k) __stcg(slot + (size_t) i * k + lane + 32 * v, T.a[i][v]);
                }
                // REC_NOP: nothing
            }
            pages_consumed += 1;
            __syncwarp();  // every lane is done reading this stage
            if (pg + kStages < n_pages) {
                if (lane == 0) issue(pg + kStages);
                pages_issued += 1;
            }
        }
        __syncwarp();
        if (lane == 0) st_release_gpu(a.done + task, epoch);
    }
}

}  // namespace sweep

#define CHB_SWEEP(NAME, T, NV, UP, MINB)                                                                          \
    extern "C" __global__ void __launch_bounds__(sweep::kWarpsPerCta * 32, MINB) NAME(const chb::SweepArgs<T> a) { \
        sweep::sweep_worker<T, NV, UP>(a);                                                                        \
    }

CHB_SWEEP(chb_sweep_f32_1_lower, float, 1, false, 4)
CHB_SWEEP(chb_sweep_f32_2_lower, float, 2, false, 4)
CHB_SWEEP(chb_sweep_f32_3_lower, float, 3, false, 3)
CHB_SWEEP(chb_sweep_f32_4_lower, float, 4, false, 3)
CHB_SWEEP(chb_sweep_f32_1_upper, float, 1, true, 4)
CHB_SWEEP(chb_sweep_f32_2_upper, float, 2, true, 4)
CHB_SWEEP(chb_sweep_f32_3_upper, float, 3, true, 3)
CHB_SWEEP(chb_sweep_f32_4_upper, float, 4, true, 3)
CHB_SWEEP(chb_sweep_f64_1_lower, double, 1, false, 3)
CHB_SWEEP(chb_sweep_f64_2_lower, double, 2, false, 2)
CHB_SWEEP(chb_sweep_f64_3_lower, double, 3, false, 2)
CHB_SWEEP(chb_sweep_f64_4_lower, double, 4, false, 2)
CHB_SWEEP(chb_sweep_f64_1_upper, double, 1, true, 3)
CHB_SWEEP(chb_sweep_f64_2_upper, double, 2, true, 2)
CHB_SWEEP(chb_sweep_f64_3_upper, double, 3, true, 2)
CHB_SWEEP(chb_sweep_f64_4_upper, double, 4, true, 2)
