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
//   * consecutive block records are processed in BATCHES: lane i fetches the ready flag of the
//     i-th block of the batch (one L2 round trip for up to 32 flags), then the source rows of every
//     ready block are copied into a per-warp shared-memory row ring with cp.async.cg (one more
//     round trip for up to kRowSlots rows, none of them through registers), then the arithmetic
//     runs out of shared memory: factor values broadcast (LDS.128 for the dense block layout),
//     source rows read conflict-free;
//   * dependencies: one epoch-valued ready flag per TASK (ld.acquire.gpu / st.release.gpu); flags
//     are never cleared (the epoch is the solve counter);
//   * the work vector `tmp` has its row stride padded to 4 elements so every row is a whole number
//     of 16-byte cp.async chunks;
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
                  REC_BLOCK_TM, REC_END_FINAL, REC_END_SLOT, REC_END_DENSE };

__host__ __device__ constexpr int row_slots(int) { return 16; }  // rows staged per batch (one full block)

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
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
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

// One staged source row (shared memory) -> this lane's columns.
template <typename F, int NV>
__device__ __forceinline__ void read_row(F (&x)[NV], const F *row, int lane, int k) {
#pragma unroll
    for (int v = 0; v < NV; ++v) x[v] = lane + 32 * v < k ? row[lane + 32 * v] : F(0);
}

// Substitution with the panel's triangle given as a dense WP x WP matrix with reciprocal pivots
// (REC_END_DENSE): fully static, one row of WP values per pivot from shared memory.
template <typename F, int NV, bool UPPER, int WP>
__device__ __forceinline__ void tri_dense(Tile<F, NV> &T, const F *m) {
#pragma unroll
    for (int step = 0; step < WP; ++step) {
        const int s = UPPER ? WP - 1 - step : step;
        F row[WP];
        if constexpr (WP == 1) {
            row[0] = m[0];
        } else {
#pragma unroll
            for (int q = 0; q < WP / 4; ++q) {
                const Vec4<F> vv = Vec4<F>::load(m + s * WP + 4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) row[4 * q + j] = vv.x[j];
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) T.a[s][v] = T.a[s][v] * row[s];
#pragma unroll
        for (int t = UPPER ? 0 : s + 1; t < (UPPER ? s : WP); ++t)
#pragma unroll
            for (int v = 0; v < NV; ++v) T.a[t][v] = fnma(row[t], T.a[s][v], T.a[t][v]);
    }
}

__device__ __forceinline__ bool is_block(uint32_t type) { return type >= REC_BLOCK_D4 && type <= REC_BLOCK_TM; }
__device__ __forceinline__ int next_source(uint32_t &mask, bool upper) {
    const int s = upper ? 31 - __clz(mask) : __ffs(mask) - 1;
    mask &= ~(1u << s);
    return s;
}

// ---- the worker ----------------------------------------------------------------------------------
// Dynamic shared memory per CTA: [warp][kStages][PAGE] slab rings, [warp][kRowSlots][kp] row rings,
// [warp][kStages] mbarriers; kp = k rounded up to 4 elements.
template <typename F, int NV, bool UPPER>
__device__ __forceinline__ void sweep_worker(const SweepArgs<F> &a) {
    constexpr int PAGE = PageOf<F>::bytes;
    constexpr int kRowSlots = row_slots((int) sizeof(F));
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int k = a.k;
    const int kp = (k + 3) & ~3;   // row stride of tmp and of the staged rows
    const int epoch = a.epoch;
    unsigned char *ring = smem_raw + (size_t) warp * kStages * PAGE;
    F *rowring = reinterpret_cast<F *>(smem_raw + (size_t) kWarpsPerCta * kStages * PAGE) + (size_t) warp * kRowSlots * kp;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t) kWarpsPerCta * kStages * PAGE +
                                                  (size_t) kWarpsPerCta * kRowSlots * kp * sizeof(F)) + warp * kStages;
    const int chunks_per_row = kp * (int) sizeof(F) / 16;

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
    long long wait_cycles = 0;  // tracing only
    int cur_task = 0;

    // Blocks until the task behind a wait code has published (acquire).
    auto wait_for = [&](uint32_t code) {
        if (code == 0) return;
        const int *flag = a.done + (code - 1);
        const long long c0 = a.trace ? clock64() : 0;
        while (ld_acquire_gpu(flag) != epoch) {}
        if (a.trace) {
            wait_cycles += clock64() - c0;
            if (lane == 0) a.trace[8 * (size_t) cur_task + 4] = global_ns();  // last blocking wait returned
        }
    };

    for (;;) {
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(a.ticket, 1);
        ticket = __shfl_sync(kFullMask, ticket, 0);
        if (ticket >= a.n_tasks) break;
        const int task = a.order[ticket];
        cur_task = task;
        const SweepTask td = a.tasks[task];
        if (a.trace && lane == 0) {
            a.trace[8 * (size_t) task] = global_ns();
            wait_cycles = 0;
        }
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
        auto store_tile = [&](const int *perm) {
            if (a.trace && lane == 0) a.trace[8 * (size_t) task + 6] = global_ns();  // triangle done
#pragma unroll
            for (int i = 0; i < W; ++i) {
                if (i < w) {
                    F *row = a.tmp + (size_t) (row0 + i) * kp;
                    F *out = UPPER ? a.dst + (size_t) perm[i] * k : nullptr;
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        if (lane + 32 * v < k) {
                            __stcg(row + lane + 32 * v, T.a[i][v]);
                            if (UPPER) out[lane + 32 * v] = T.a[i][v];
                        }
                }
            }
        };

        // One block record, its source rows already staged at `xrow` (consecutive slots, sweep order).
        auto run_block = [&](uint32_t type, const uint4 &h, const unsigned char *pay, const F *xrow) {
            const uint32_t tmask = h.z >> 16, smask = h.z & 0xffff;
            const F *val = reinterpret_cast<const F *>(pay);
            const int ns = __popc(smask);
            if (type == REC_BLOCK_D4) {
                const int stride = (32 - __clz(tmask) + 3) & ~3;
                for (int r = 0; r < ns; ++r) {
                    F x[NV];
                    read_row<F, NV>(x, xrow + (size_t) r * kp, lane, k);
                    const F *vr = val + r * stride;
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
                }
            } else if (type == REC_BLOCK_D1) {
                const int nt = __popc(tmask);
                for (int r = 0; r < ns; ++r) {
                    F x[NV];
                    read_row<F, NV>(x, xrow + (size_t) r * kp, lane, k);
                    const F *vr = val + r * nt;
#pragma unroll
                    for (int t = 0; t < 3; ++t)
                        if (t < nt) {
                            const F l = vr[t];
#pragma unroll
                            for (int v = 0; v < NV; ++v) T.a[t][v] = fnma(l, x[v], T.a[t][v]);
                        }
                }
            } else {  // REC_BLOCK_TM: target-major, arbitrary target set
                uint32_t tm = tmask;
                while (tm) {
                    const int t = __ffs(tm) - 1;
                    tm &= tm - 1;
                    F acc[NV];
                    get_row<F, NV>(T, t, acc);
                    for (int r = 0; r < ns; ++r) {
                        F x[NV];
                        read_row<F, NV>(x, xrow + (size_t) r * kp, lane, k);
                        const F l = val[r];
#pragma unroll
                        for (int v = 0; v < NV; ++v) acc[v] = fnma(l, x[v], acc[v]);
                    }
                    val += ns;
                    put_row<F, NV>(T, t, acc);
                }
            }
        };

        for (uint32_t pg = 0; pg < n_pages; ++pg) {
            const uint32_t stage = pages_consumed % kStages;
            mbar_wait(bars + stage, (pages_consumed / kStages) & 1);
            if (a.trace && lane == 0 && pg == 0) a.trace[8 * (size_t) task + 1] = global_ns();
            const unsigned char *p = ring + (size_t) stage * PAGE;
            const unsigned char *pend = p + min((uint32_t) PAGE, total - pg * PAGE);
            while (p < pend) {
                const uint4 h = *reinterpret_cast<const uint4 *>(p);
                const uint32_t type = h.x & 0xff;
                const uint32_t len = (h.x >> 8) * 16;
                const unsigned char *pay = p + 16;
                if (is_block(type)) {
                    // ---- batch: consecutive block records whose source rows fit the row ring ----
                    int n_blocks = 0, n_rows = 0;
                    uint32_t my_code = 0;
                    for (const unsigned char *q = p; q < pend && n_blocks < 32;) {
                        const uint4 hq = *reinterpret_cast<const uint4 *>(q);
                        if (!is_block(hq.x & 0xff)) break;
                        const int ns = __popc(hq.z & 0xffff);
                        if (n_rows + ns > kRowSlots) break;
                        if (lane == n_blocks) my_code = hq.w;
                        n_rows += ns;
                        n_blocks += 1;
                        q += (hq.x >> 8) * 16;
                    }
                    // one ready flag per lane; blocks before the first unfinished dependency go now
                    bool ready = true;
                    if (my_code != 0) ready = ld_acquire_gpu(a.done + (my_code - 1)) == epoch;
                    const unsigned pending = __ballot_sync(kFullMask, !ready);
                    int n_go = pending ? __ffs(pending) - 1 : n_blocks;
                    if (n_go == 0) {
                        wait_for(h.w);
                        n_go = 1;
                    }
                    // stage the source rows of those blocks
                    {
                        const unsigned char *q = p;
                        int slot = 0;
                        for (int i = 0; i < n_go; ++i) {
                            const uint4 hq = *reinterpret_cast<const uint4 *>(q);
                            uint32_t sm = hq.z & 0xffff;
                            const F *rows = a.tmp + (size_t) (int) hq.y * kp;
                            while (sm) {
                                const int s = next_source(sm, UPPER);
                                const unsigned char *src = reinterpret_cast<const unsigned char *>(rows + (size_t) s * kp);
                                unsigned char *dst = reinterpret_cast<unsigned char *>(rowring + (size_t) slot * kp);
                                for (int c = lane; c < chunks_per_row; c += 32) cp_async_16(dst + 16 * c, src + 16 * c);
                                slot += 1;
                            }
                            q += (hq.x >> 8) * 16;
                        }
                        cp_async_wait_all();
                        __syncwarp();
                    }
                    // arithmetic out of shared memory
                    {
                        int slot = 0;
                        for (int i = 0; i < n_go; ++i) {
                            const uint4 hq = *reinterpret_cast<const uint4 *>(p);
                            run_block(hq.x & 0xff, hq, p + 16, rowring + (size_t) slot * kp);
                            slot += __popc(hq.z & 0xffff);
                            p += (hq.x >> 8) * 16;
                        }
                    }
                    __syncwarp();  // the row ring is reused by the next batch
                    continue;
                }
                p += len;
                if (type == REC_BEGIN_RHS) {
                    row0 = (int) h.y;
                    w = (int) h.z;
                    const int *perm = reinterpret_cast<const int *>(pay);
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        const F *row = nullptr;
                        if (i < w) row = UPPER ? a.tmp + (size_t) (row0 + i) * kp : a.src + (size_t) perm[i] * k;
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            F val = F(0);
                            if (i < w && lane + 32 * v < k) val = UPPER ? __ldcg(row + lane + 32 * v) : __ldg(row + lane + 32 * v);
                            T.a[i][v] = val;
                        }
                    }
                } else if (type == REC_END_DENSE) {
                    if (a.trace && lane == 0) a.trace[8 * (size_t) task + 5] = global_ns();  // blocks done
                    const int *perm = reinterpret_cast<const int *>(pay);
                    const F *m = reinterpret_cast<const F *>(pay + (UPPER ? 4 * W : 0));
                    if (w == 1) tri_dense<F, NV, UPPER, 1>(T, m);
                    else if (w <= 4) tri_dense<F, NV, UPPER, 4>(T, m);
                    else if (w <= 8) tri_dense<F, NV, UPPER, 8>(T, m);
                    else if (w <= 12) tri_dense<F, NV, UPPER, 12>(T, m);
                    else tri_dense<F, NV, UPPER, 16>(T, m);
                    store_tile(perm);
                } else if (type == REC_END_FINAL) {
                    // reference-order substitution: packed triangle, IEEE division (exact_order schedules)
                    if (a.trace && lane == 0) a.trace[8 * (size_t) task + 5] = global_ns();
                    const int *perm = reinterpret_cast<const int *>(pay);
                    const F *tv = reinterpret_cast<const F *>(pay + (UPPER ? 4 * W : 0));
#pragma unroll 1
                    for (int step = 0; step < w; ++step) {
                        const int s = UPPER ? w - 1 - step : step;
                        const F d = *tv++;
                        F x[NV];
                        get_row<F, NV>(T, s, x);
#pragma unroll
                        for (int v = 0; v < NV; ++v) x[v] = x[v] / d;
                        put_row<F, NV>(T, s, x);
#pragma unroll
                        for (int t = 0; t < W; ++t) {
                            if (UPPER ? t < s : (t > s && t < w)) {
                                const F l = *tv++;
#pragma unroll
                                for (int v = 0; v < NV; ++v) T.a[t][v] = fnma(l, x[v], T.a[t][v]);
                            }
                        }
                    }
                    store_tile(perm);
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
        if (a.trace && lane == 0) {
            a.trace[8 * (size_t) task + 2] = global_ns();
            a.trace[8 * (size_t) task + 3] = (unsigned long long) wait_cycles;
        }
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
