// Panel sweep kernels (sm_100a): the forward (L y = P b) and backward (L' x = y, x -> P' x) triangular
// solves over the panel / block / task image built by panel_schedule.cpp.
//
// What they replace: solve_lower_* / solve_upper_* of the reference (kernels/kernels.cu:64-131,
// 133-202: one 128-thread CTA per ROW, one flag poll + two block barriers per non-zero).
//
// Execution model
//   * persistent, co-resident grid; every CTA is one worker: 1-4 COMPUTE warps, warp w owning RHS
//     columns [32 w, 32 w + 32), plus one LOADER warp.  All compute warps walk the same task's
//     record stream, so a factor byte is fetched once per CTA and used for up to 128 columns, and a
//     task's arithmetic is spread over up to four warp schedulers (the task graph's critical path
//     does not grow with n_rhs);
//   * the loader draws tickets from one atomic counter; ticket order is a topological order of the
//     task graph, so the oldest unfinished ticket is always running with all its dependencies done
//     (forward progress).  It streams each task's slab (record headers + factor values, one
//     contiguous byte range) into a ring of pages with 1-D TMA bulk copies (cp.async.bulk ...
//     mbarrier::complete_tx; factor bytes leave HBM exactly once per sweep), parses the records
//     ahead of the compute warps, and for every record that consumes already-solved rows it polls
//     the producer's ready flag (32 flags per round trip, one per lane) and copies the rows from the
//     work vector into a shared-memory ROW RING, again with TMA bulk copies completing on that
//     record's mbarrier.  It runs ahead by the depth of the rings, across task boundaries, so flag
//     and gather latency overlap the arithmetic of earlier records;
//   * compute warps never poll global flags on this path: they wait on the record's shared-memory
//     mbarrier, multiply out of shared memory (factor values broadcast with LDS.128, source rows
//     conflict-free), and hand the ring rows back through an "empty" mbarrier;
//   * the tile (<= 16 rows) of the panel being solved lives in registers, one RHS column per lane;
//   * LOCAL tasks (complete subtrees of small panels, most of the rows) are executed by their own
//     kernel by default (local_sweep.cuh); this kernel keeps the code for them (CHB_LOCAL_KERNEL=0, or
//     tasks over that kernel's budget): the subtree's right-hand side rows and the external rows it
//     reads are staged into the row ring by the loader as ONE record, every record of the task then
//     works in place in shared memory, and the solved rows are written back once;
//   * dependencies: one epoch-valued ready flag per TASK (ld.acquire.gpu / st.release.gpu after a
//     named barrier of the compute warps); flags are never cleared (the epoch is the solve counter);
//   * the work vector `tmp` and the partial-tile scratch have their row stride padded to 4 elements
//     so every row is a whole number of 16-byte units (TMA granularity); a record's rows never wrap
//     around the end of the row ring (the loader skips a too-short tail), so rows are addressed
//     without modular arithmetic;
//   * tickets are drawn one at a time near narrow dependency levels and eight at a time deep inside
//     wide ones (host hint in the task table), and the head of the next batch (descriptor + first
//     slab page) is fetched while the compute warps finish the current one;
//   * P is fused into the forward tile load, P' into the backward tile store;
//   * instruction footprint is a first-order concern (L0 holds ~6 KB of SASS, L1.5 32 KB): one block
//     routine, one triangle routine, loops instead of unrolled variants, exact-order code in
//     separate kernels.  Measured: every ~100 SASS instructions removed bought about 1 %.
#pragma once
#include <cstdint>

#include "kernel_args.h"

namespace sweep {

using chb::SweepArgs;
using chb::SweepTask;

constexpr int W = 16;            // panel width == tile rows
constexpr int kStages = 4;       // pages in the CTA's slab ring
constexpr int kPage = 4096;      // bytes per page (panel_schedule.h: page_bytes_for)
constexpr int kRowBars = 16;     // staged records in flight (full/empty mbarrier pairs of the row ring)
constexpr int kTaskQueue = 8;    // task descriptors the loader may run ahead of the compute warps
constexpr int kBatch = 8;        // tickets drawn at a time by a loader (a power of two <= 32)
constexpr unsigned kFullMask = 0xffffffffu;

enum : uint32_t { REC_NOP = 0, REC_BEGIN_RHS, REC_BEGIN_ZERO, REC_BEGIN_SLOT, REC_ADD_SLOT, REC_BLOCK_D4, REC_BLOCK_D1,
                  REC_BLOCK_TM, REC_END_FINAL, REC_END_SLOT, REC_END_DENSE, REC_LOCAL_LOAD, REC_LOCAL_STORE,
                  REC_END_STORE, REC_LOCAL_ROWS, REC_ADD_ROWS, REC_LOCAL_FLAG = 0x40, REC_SRC_R_FLAG = 0x80 };

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A waiting warp is suspended by the hardware (up to the hint) instead of spinning through the issue slots of the
// working warps on its scheduler: spin instructions were 12 % of everything issued on cfg4 before (ncu).
constexpr uint32_t kSuspendHintNs = 20000;
__device__ __forceinline__ unsigned long long global_ns();
// Watchdog of every wait: 20 s of %globaltimer, not an attempt count (the suspend hint is an upper bound, so a count
// is a budget of anything between milliseconds and minutes: a debugger, compute-sanitizer or a pre-empted context
// must not trip it).  A dependency that never arrives is a schedule or kernel bug: trap rather than hang the GPU.
constexpr unsigned long long kWatchdogNs = 20ull * 1000ull * 1000ull * 1000ull;
#ifndef CHB_AB_COMPUTE_BACKOFF
#define CHB_AB_COMPUTE_BACKOFF 0
#endif
// BACKOFF_NS > 0: a failed attempt is followed by a real sleep (compute warps; CHB_AB_COMPUTE_BACKOFF experiments)
template <int BACKOFF_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        uint32_t ok;
#ifdef CHB_AB_NO_HINT
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
#else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs)
            : "memory");
#endif
        if (ok) return;
        if (BACKOFF_NS > 0) __nanosleep(BACKOFF_NS);
#ifdef CHB_AB_COUNT_WATCHDOG
        if (++spins > (1u << 22)) __trap();
#else
        if ((++spins & 255u) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kWatchdogNs) __trap();
        }
#endif
    }
}
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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
// T[t] for a warp-uniform, run-time t (registers cannot be indexed dynamically: jump table)
#define SWEEP_ROW_CASES(OP) \
    switch (t) {            \
        case 0: OP(0) break; case 1: OP(1) break; case 2: OP(2) break; case 3: OP(3) break;     \
        case 4: OP(4) break; case 5: OP(5) break; case 6: OP(6) break; case 7: OP(7) break;     \
        case 8: OP(8) break; case 9: OP(9) break; case 10: OP(10) break; case 11: OP(11) break; \
        case 12: OP(12) break; case 13: OP(13) break; case 14: OP(14) break; default: OP(15) break; \
    }

template <typename F> __device__ __forceinline__ F get_row(const F (&T)[W], int t) {
    F r;
#define SWEEP_GET(I) { r = T[I]; }
    SWEEP_ROW_CASES(SWEEP_GET)
#undef SWEEP_GET
    return r;
}
template <typename F> __device__ __forceinline__ void put_row(F (&T)[W], int t, F r) {
#define SWEEP_PUT(I) { T[I] = r; }
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

// Substitution with the panel's triangle given as a dense wp x wp matrix (wp = 1 or a multiple of 4) with
// reciprocal pivots (REC_END_DENSE).  One copy of the code for every width (the instruction footprint
// of these kernels matters: L0 holds ~6 KB of SASS, L1.5 32 KB): the pivot loop is unrolled over all
// 16 steps so the tile stays in registers, steps and 4-value groups beyond wp are skipped by
// warp-uniform branches.
template <typename F, bool UPPER>
__device__ __forceinline__ void tri_dense(F (&T)[W], const F *m, int wp) {
#pragma unroll
    for (int step = 0; step < W; ++step) {
        const int s = UPPER ? W - 1 - step : step;
        if (s < wp) {
            const F *row = m + s * wp;
            T[s] = T[s] * row[s];
#pragma unroll
            for (int q = 0; q < W / 4; ++q) {
                const bool needed = UPPER ? 4 * q < s : 4 * q + 3 > s;  // the group holds an unknown still to update
                if (needed && 4 * q < wp) {
                    const Vec4<F> vv = Vec4<F>::load(row + 4 * q);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int t = 4 * q + j;
                        if (UPPER ? t < s : t > s) T[t] = fnma(vv.x[j], T[s], T[t]);
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// Spin loops trap instead of hanging the GPU if a dependency never arrives (a schedule or kernel bug).
#define SWEEP_SPIN_GUARD(counter)                                                          \
    if ((++(counter) & 0xffffu) == 0) {                                                    \
        const unsigned long long guard_now = global_ns();                                  \
        if (guard_t0 == 0) guard_t0 = guard_now;                                           \
        if (guard_now - guard_t0 > kWatchdogNs) __trap();                                  \
    }

// A task descriptor in the CTA's shared-memory task queue (loader -> compute warps).
struct QueuedTask {
    int ticket;   // >= n_tasks: no more work
    int task;     // id = index of the task's ready flag
    unsigned int bytes;
    unsigned int pad;
    unsigned long long offset;
    unsigned long long pad2;
};

// One block record: tile -= L_block * sources ("grouped source-major" values, panel_schedule.h).  `ring`
// points at this lane's column of ring row 0 (rows are kp elements apart, the ring has ring_rows rows).
// The record's ns source rows start at ring row `base`, in ASCENDING row order (the backward sweep visits them
// descending).  A plain loop on purpose (small instruction footprint); two sources per iteration hide the LDS latency.
template <typename F, bool UPPER>
__device__ __forceinline__ void run_block(F (&T)[W], uint32_t masks, const unsigned char *pay, const F *ring, int kp, int base) {
    const uint32_t tmask = masks >> 16;
    const int ns = __popc(masks & 0xffff);
    const bool g0 = (tmask & 0x000f) != 0, g1 = (tmask & 0x00f0) != 0, g2 = (tmask & 0x0f00) != 0, g3 = (tmask & 0xf000) != 0;
    const int stride = 4 * ((int) g0 + (int) g1 + (int) g2 + (int) g3);
    const F *val = reinterpret_cast<const F *>(pay);
    // a record's rows are contiguous in the ring (the loader never lets one wrap around)
    const F *src = ring + (size_t) (UPPER ? base + ns - 1 : base) * kp;
#pragma unroll 2
    for (int r = 0; r < ns; ++r) {
        const F x = *src;
        src = UPPER ? src - kp : src + kp;
        const F *v = val;
        if (g0) {
            const Vec4<F> vv = Vec4<F>::load(v);
            v += 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) T[j] = fnma(vv.x[j], x, T[j]);
        }
        if (g1) {
            const Vec4<F> vv = Vec4<F>::load(v);
            v += 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) T[4 + j] = fnma(vv.x[j], x, T[4 + j]);
        }
        if (g2) {
            const Vec4<F> vv = Vec4<F>::load(v);
            v += 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) T[8 + j] = fnma(vv.x[j], x, T[8 + j]);
        }
        if (g3) {
            const Vec4<F> vv = Vec4<F>::load(v);
#pragma unroll
            for (int j = 0; j < 4; ++j) T[12 + j] = fnma(vv.x[j], x, T[12 + j]);
        }
        val += stride;
    }
}

// Records whose input rows the LOADER stages into the row ring.  Returns the number of rows, 0 if the
// record needs no rows (those are handled by the compute warps alone).
__device__ __forceinline__ int staged_rows(const uint4 &h) {
    const uint32_t type = h.x & 0x7f;  // records inside local tasks carry REC_LOCAL_FLAG and never match
    if (type == REC_BLOCK_D4) return __popc(h.z & 0xffff);
    if (type == REC_BEGIN_SLOT || type == REC_ADD_SLOT) return W;
    if (type == REC_ADD_ROWS) return (int) (h.z & 0xffff);
    if (type == REC_BEGIN_RHS) return (int) h.z;
    if (type == REC_LOCAL_LOAD) return (int) (h.z & 0xffff) + (int) (h.z >> 16);
    return 0;
}

// Shared memory of one CTA (dynamic):
//   [kStages][kPage]          slab ring (TMA destination)
//   [ring_rows][kp]           row ring: right-hand-side rows, already-solved rows, partial tiles and the
//                             rows of local tasks, all staged by the loader
//   full[kStages], empty[kStages]      slab ring barriers   (empty: compute warps + loader)
//   rfull[kRowBars], rempty[kRowBars]  row ring barriers    (rempty: compute warps)
//   tfull[kTaskQueue], tempty[kTaskQueue]
//   QueuedTask[kTaskQueue], int rows_of[kRowBars]
template <typename F>
struct Smem {
    unsigned char *slab;
    F *ring;
    uint64_t *full, *empty, *rfull, *rempty, *tfull, *tempty;
    QueuedTask *queue;
    int *rows_of;
    __device__ __forceinline__ Smem(unsigned char *raw, int kp, int ring_rows) {
        slab = raw;
        ring = reinterpret_cast<F *>(raw + (size_t) kStages * kPage);
        full = reinterpret_cast<uint64_t *>(ring + (size_t) ring_rows * kp);
        empty = full + kStages;
        rfull = empty + kStages;
        rempty = rfull + kRowBars;
        tfull = rempty + kRowBars;
        tempty = tfull + kTaskQueue;
        queue = reinterpret_cast<QueuedTask *>(tempty + kTaskQueue);
        rows_of = reinterpret_cast<int *>(queue + kTaskQueue);
    }
};

// ---- the loader warp -------------------------------------------------------------------------------
// Draws tickets, streams each task's slab into the page ring, and for every staged record (in order)
// polls the ready flag(s) of its producer(s) and copies its input rows into the row ring with 1-D TMA
// bulk copies completing on that record's mbarrier (runs of consecutive rows are one copy).  It is the
// only warp that reads the right-hand side, the work vector or the scratch tiles from global memory,
// and it runs ahead of the compute warps by up to the depth of the three rings, across task boundaries.
template <typename F, bool UPPER>
__device__ __forceinline__ void loader_warp(const SweepArgs<F> &a, const Smem<F> &sm, int lane) {
    const int k = a.k, kp = (k + 3) & ~3;
    const int ring_rows = a.ring_rows;
    const uint32_t row_bytes = (uint32_t) kp * (uint32_t) sizeof(F);
    const uint32_t src_row_bytes = (uint32_t) k * (uint32_t) sizeof(F);
    const bool src_tma = a.src_tma != 0;  // rows of b are 16-byte aligned multiples of 16 bytes
    const int epoch = a.epoch;
    uint32_t gpage = 0;       // global index of the current task's first page
    uint32_t issued = 0;      // pages handed to TMA so far (lane 0 only)
    uint32_t j_issue = 0;     // staged records issued so far == sequence number of the next one
    uint32_t j_rel = 0;       // staged records known to be released by the compute warps
    int rows_in_flight = 0;   // ring rows of records j_rel .. j_issue - 1
    int head = 0;             // ring row the next record starts at

    // n consecutive rows (row_bytes each, contiguous in global memory) -> ring rows slot, slot + 1, ...
    auto copy_rows = [&](int slot, const F *src, int n, uint64_t *bar) {
        tma_load_1d(sm.ring + (size_t) slot * kp, src, (uint32_t) n * row_bytes, bar);
    };
    // rows b[perm[i]], i < n -> ring rows head + i (forward sweep only); returns the TMA bytes to expect
    auto copy_rhs_rows = [&](const int *perm, int n) -> uint32_t {
        if (src_tma) {
            for (int i = lane; i < n; i += 32)
                tma_load_1d(sm.ring + (size_t) (head + i) * kp, a.src + (size_t) perm[i] * a.src_ld, src_row_bytes,
                            sm.rfull + j_issue % kRowBars);
            return (uint32_t) n * src_row_bytes;
        }
        const int total = n * k;
#pragma unroll 1
        for (int e = lane; e < total; e += 32) {
            const int i = e / k, c = e - i * k;
            sm.ring[(size_t) (head + i) * kp + c] = __ldg(a.src + (size_t) perm[i] * a.src_ld + c);
        }
        return 0u;
    };

    // Tickets are drawn kBatch at a time (consecutive, processed in order: forward progress is unaffected), so the
    // descriptor fetch costs one round trip per batch and the slab pages of the batch's later tasks are
    // already in flight while its earlier tasks are staged: per-task latency of the loader is what bounds
    // the throughput of small tasks.
    uint32_t q = 0;            // tasks queued so far
    int t0 = 0;                // first ticket of the current batch
    SweepTask td_head{0ull, 0u, 0u};  // its descriptor (lane 0), fetched at the end of the previous batch
    if (lane == 0) {
        t0 = atomicAdd(a.ticket, 1);
        if (t0 < a.n_tasks) td_head = a.tasks[t0];
    }
    for (;;) {
        t0 = __shfl_sync(kFullMask, t0, 0);
        if (t0 >= a.n_tasks) {
            if (lane == 0) {
                if (q >= kTaskQueue) mbar_wait(sm.tempty + q % kTaskQueue, ((q / kTaskQueue) - 1) & 1);
                sm.queue[q % kTaskQueue] = QueuedTask{a.n_tasks, 0, 0u, 0u, 0ull, 0ull};
                mbar_arrive(sm.tfull + q % kTaskQueue);
            }
            break;
        }
        // The descriptor of the first task says how many tickets may be drawn together here (bits 28-30 of
        // `bytes`, set by the host: > 1 only deep inside a wide dependency level, where every worker is busy with
        // tasks of that level anyway; near narrow levels one ticket at a time keeps the critical path short).
        SweepTask td = td_head;
        const int hint = (int) ((__shfl_sync(kFullMask, td.bytes, 0) >> 28) & 7u) + 1;
        int t1 = 0;
        if (hint > 1) {
            if (lane == 0) t1 = atomicAdd(a.ticket, hint - 1);
            t1 = __shfl_sync(kFullMask, t1, 0);
        }
        const int my_ticket = lane == 0 ? t0 : t1 + lane - 1;
        const bool valid = lane < hint && my_ticket < a.n_tasks;
        if (lane > 0 && valid) td = a.tasks[my_ticket];
        td.bytes &= 0x0fffffffu;
        if (!valid) td.bytes = 0;
        const int n_batch = __popc(__ballot_sync(kFullMask, valid));
        int t_ahead = a.n_tasks;
        if (lane == 0) t_ahead = atomicAdd(a.ticket, 1);  // first ticket of the next batch: in flight during this one
        const uint32_t my_pages = (td.bytes + kPage - 1) / kPage;
        uint32_t my_first = my_pages;  // inclusive scan ...
#pragma unroll
        for (int o = 1; o < kBatch; o <<= 1) {
            const uint32_t up = __shfl_up_sync(kFullMask, my_first, o);
            if (lane >= o) my_first += up;
        }
        const uint32_t batch_pages = __shfl_sync(kFullMask, my_first, kBatch - 1);
        my_first = gpage + my_first - my_pages;  // ... made exclusive and global
        const uint32_t batch_end = gpage + batch_pages;
        int iss_task = 0;  // batch task owning page `issued` (warp-uniform)
        // hands pages to TMA in global order while the ring has room: up to page g + kStages - 1; blocks only for page g
        auto feed_pages = [&](uint32_t g) {
            const uint32_t limit = min(batch_end, g + kStages);
            while (issued < limit) {
                while (issued >= __shfl_sync(kFullMask, my_first + my_pages, iss_task)) iss_task += 1;
                const uint32_t st = issued % kStages, fill = issued / kStages;
                if (fill > 0) {
                    if (issued == g) {
                        mbar_wait(sm.empty + st, (fill - 1) & 1);
                    } else {
                        const int free_now = lane == 0 ? (int) mbar_test(sm.empty + st, (fill - 1) & 1) : 0;
                        if (!__shfl_sync(kFullMask, free_now, 0)) break;
                    }
                }
                const uint32_t first = __shfl_sync(kFullMask, my_first, iss_task);
                const uint32_t tbytes = __shfl_sync(kFullMask, td.bytes, iss_task);
                const uint32_t lo = __shfl_sync(kFullMask, (uint32_t) td.offset, iss_task);
                const uint32_t hi = __shfl_sync(kFullMask, (uint32_t) (td.offset >> 32), iss_task);
                const uint32_t pi = issued - first;
                const uint32_t bytes = min((uint32_t) kPage, tbytes - pi * kPage);
                if (lane == 0) {
                    mbar_expect_tx(sm.full + st, bytes);
                    tma_load_1d(sm.slab + (size_t) st * kPage,
                                a.slab + ((unsigned long long) lo | (unsigned long long) hi << 32) + (size_t) pi * kPage, bytes, sm.full + st);
                }
                issued += 1;
            }
        };
      for (int bi = 0; bi < n_batch; ++bi, ++q) {
        // ---- task descriptor -> queue ---------------------------------------------------------------
        const int task = (int) __shfl_sync(kFullMask, td.level, bi);
        const uint32_t total = __shfl_sync(kFullMask, td.bytes, bi);
        const uint32_t off_lo = __shfl_sync(kFullMask, (uint32_t) td.offset, bi);
        const uint32_t off_hi = __shfl_sync(kFullMask, (uint32_t) (td.offset >> 32), bi);
        const uint32_t task_first = __shfl_sync(kFullMask, my_first, bi);
        const int ticket = __shfl_sync(kFullMask, my_ticket, bi);
        if (lane == 0) {
            if (q >= kTaskQueue) mbar_wait(sm.tempty + q % kTaskQueue, ((q / kTaskQueue) - 1) & 1);
            sm.queue[q % kTaskQueue] = QueuedTask{ticket, task, total, 0u, (unsigned long long) off_lo | (unsigned long long) off_hi << 32, 0ull};
            mbar_arrive(sm.tfull + q % kTaskQueue);
        }
        const uint32_t n_pages = (total + kPage - 1) / kPage;

        for (uint32_t pg = 0; pg < n_pages; ++pg) {
            const uint32_t g = task_first + pg, stage = g % kStages;
            feed_pages(g);
            __syncwarp();
            mbar_wait(sm.full + stage, (g / kStages) & 1);
            const unsigned char *p = sm.slab + (size_t) stage * kPage;
            const unsigned char *pend = p + min((uint32_t) kPage, total - pg * kPage);
            while (p < pend) {
                // scan: the next (up to 32) staged records of this page; lane i keeps record i.  A record with
                // several producers (the backward LOCAL_LOAD) is handled alone.
                int n = 0;
                uint32_t my_off = 0, my_code = 0;
                bool multi = false;
                const unsigned char *qp = p;
                while (qp < pend && n < 32) {
                    const uint4 hq = *reinterpret_cast<const uint4 *>(qp);
                    if (staged_rows(hq) > 0) {
                        const uint32_t tq = hq.x & 0x7f;
                        const bool many = tq == REC_LOCAL_LOAD && hq.w != 0;
                        if (many && n > 0) break;
                        if (lane == n) {
                            my_off = (uint32_t) (qp - p);
                            my_code = (tq == REC_BEGIN_RHS || tq == REC_LOCAL_LOAD) ? 0u : hq.w;
                        }
                        n += 1;
                        if (many) {
                            multi = true;
                            break;
                        }
                    }
                    qp += (hq.x >> 8) * 16;
                }
                if (n == 0) break;  // nothing left to stage in this page
                int n_go = n;
                if (multi) {
                    // every producer of the local task's external rows must have published
                    const uint4 h = *reinterpret_cast<const uint4 *>(p + __shfl_sync(kFullMask, my_off, 0));
                    const int n_ext = (int) (h.z >> 16), n_wait = (int) h.w;
                    const int *waits = reinterpret_cast<const int *>(p + __shfl_sync(kFullMask, my_off, 0) + 16) + ((n_ext + 3) & ~3);
                    for (int e = lane; e < n_wait; e += 32) {
                        const int *flag = a.done + (waits[e] - 1);
                        unsigned spins = 0;
                        unsigned long long guard_t0 = 0;
                        while (ld_acquire_gpu(flag) != epoch) { SWEEP_SPIN_GUARD(spins) }
                    }
                    __syncwarp();
                } else {
                    // poll: one flag per lane, the ready prefix goes now
                    unsigned spins = 0;
                    unsigned long long guard_t0 = 0;
                    for (;;) {
                        bool ready = true;
                        if (lane < n && my_code != 0) ready = ld_acquire_gpu(a.done + (my_code - 1)) == epoch;
                        const unsigned pending = __ballot_sync(kFullMask, !ready);
                        n_go = pending ? __ffs(pending) - 1 : n;
                        if (n_go > 0) break;
                        SWEEP_SPIN_GUARD(spins)
                    }
                }
                // the producers' rows were written with generic-proxy stores and are read through the async proxy (TMA) below
#ifndef CHB_AB_NO_POLL_FENCE
                asm volatile("fence.proxy.async;" ::: "memory");
#endif
                const unsigned char *last = p;
                for (int i = 0; i < n_go; ++i) {
                    const unsigned char *rec = p + __shfl_sync(kFullMask, my_off, i);
                    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
                    const uint32_t type = h.x & 0x7f;
                    const int nr = staged_rows(h);
                    // a record's rows never wrap around the ring: skip the tail if it is too short (the skipped rows
                    // are accounted to this record).  Back-pressure: a free barrier pair and enough free rows.
                    int skipped = 0;
                    if (head + nr > ring_rows) skipped = ring_rows - head, head = 0;
                    // (The row count is conservative -- a skipped tail counts as occupied -- so with nothing in flight the
                    // record is always admitted: nr <= ring_rows by construction, and waiting could never end.)
                    while (j_issue - j_rel >= (uint32_t) kRowBars || (rows_in_flight > 0 && rows_in_flight + nr + skipped > ring_rows)) {
                        mbar_wait(sm.rempty + j_rel % kRowBars, (j_rel / kRowBars) & 1);
                        rows_in_flight -= sm.rows_of[j_rel % kRowBars];
                        j_rel += 1;
                    }
                    __syncwarp();
                    uint64_t *bar = sm.rfull + j_issue % kRowBars;
                    if (lane == 0) sm.rows_of[j_issue % kRowBars] = nr + skipped;
                    uint32_t tx = (uint32_t) nr * row_bytes;
                    // what this lane copies: n_copy consecutive rows from `from` to ring row `slot` (one TMA copy, two
                    // if it wraps); rows of the user's b (scattered through P) are handled below
                    const F *from = nullptr;
                    int slot = head, n_copy = 0;
                    const int *rhs_perm = nullptr;
                    int rhs_n = 0;
                    if (type == REC_BLOCK_D4) {
                        // runs of consecutive source rows, ascending: lane r copies run r
                        uint32_t m = h.z & 0xffff;
                        int rank = 0, n_runs = 0;
                        const F *rows = ((h.x & REC_SRC_R_FLAG) ? a.rbuf : a.tmp) + (size_t) (int) h.y * kp;
                        while (m) {
                            const int s0 = __ffs(m) - 1;
                            const int len = __ffs(~(m >> s0)) - 1;
                            if (lane == n_runs) from = rows + (size_t) s0 * kp, n_copy = len, slot = head + rank;
                            rank += len;
                            m &= ~(((1u << len) - 1u) << s0);
                            n_runs += 1;
                        }
                    } else if (type == REC_BEGIN_SLOT || type == REC_ADD_SLOT) {
                        if (lane == 0) from = a.scratch + (size_t) (h.z >> 8) * W * kp, n_copy = W;
                    } else if (!UPPER && type == REC_ADD_ROWS) {
                        if (lane == 0) from = a.cbuf + (size_t) h.y * kp, n_copy = nr;
                    } else if (type == REC_BEGIN_RHS) {
                        if (UPPER) {
                            if (lane == 0) from = a.tmp + (size_t) (int) h.y * kp, n_copy = nr;
                        } else {
                            rhs_perm = reinterpret_cast<const int *>(rec + 16), rhs_n = nr;
                        }
                    } else {  // REC_LOCAL_LOAD
                        const int n_int = (int) (h.z & 0xffff), n_ext = (int) (h.z >> 16);
                        const int *list = reinterpret_cast<const int *>(rec + 16);
                        if (UPPER) {  // lane 31: the task's own rows (one run); lanes < n_ext (<= 31): one external row each
                            if (lane == 31) from = a.tmp + (size_t) (int) h.y * kp, n_copy = n_int;
                            else if (lane < n_ext) from = a.tmp + (size_t) list[lane] * kp, n_copy = 1, slot = head + n_int + lane;
                        } else {
                            rhs_perm = list, rhs_n = n_int;
                        }
                    }
                    if (n_copy > 0) copy_rows(slot, from, n_copy, bar);
                    if (rhs_n > 0) tx = copy_rhs_rows(rhs_perm, rhs_n);
                    if (rhs_n > 0 && !src_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes to ring rows TMA reuses later
                    __syncwarp();  // manual copies (unaligned right-hand side) are complete before the arrival
                    if (lane == 0) mbar_expect_tx(bar, tx);
                    j_issue += 1;
                    rows_in_flight += nr + skipped;
                    head += nr;
                    last = rec + (h.x >> 8) * 16;
                }
                p = last;
            }
            __syncwarp();  // every lane is done reading this stage
            if (lane == 0) mbar_arrive(sm.empty + stage);
        }
      }
        gpage = batch_end;
        // Head of the next batch, fetched NOW -- while the compute warps are still busy with the tail of this batch --
        // so that neither the descriptor round trip nor the first slab page's TMA latency is exposed at its start:
        // on narrow dependency levels (one ticket per batch) that is two round trips off every hop of the chain.
        t0 = __shfl_sync(kFullMask, t_ahead, 0);
        td_head = SweepTask{0ull, 0u, 0u};
        if (t0 < a.n_tasks) {
            if (lane == 0) td_head = a.tasks[t0];
            const uint32_t st = issued % kStages, fill = issued / kStages;
            int free_now = 1;
            if (fill > 0 && lane == 0) free_now = (int) mbar_test(sm.empty + st, (fill - 1) & 1);
            if (issued == gpage && __shfl_sync(kFullMask, free_now, 0)) {  // nothing of the next batch is in flight yet
                if (lane == 0) {
                    const uint32_t bytes = min((uint32_t) kPage, td_head.bytes & 0x0fffffffu);
                    mbar_expect_tx(sm.full + st, bytes);
                    tma_load_1d(sm.slab + (size_t) st * kPage, a.slab + td_head.offset, bytes, sm.full + st);
                }
                issued += 1;
            }
        }
    }
}

// ---- the compute warps -----------------------------------------------------------------------------
template <typename F, bool UPPER, bool EXACT, bool TRACE>
__device__ __forceinline__ void compute_warp(const SweepArgs<F> &a, const Smem<F> &sm, int warp, int lane, int n_compute) {
    const int k = a.k, kp = (k + 3) & ~3;
    const int ring_rows = a.ring_rows;
    const int col = warp * 32 + lane;
    const bool active = col < k;
    const int lcol = active ? col : warp * 32;   // a column this lane may always read
    F *ring = sm.ring + lcol;                    // this lane's column of ring row 0
    const bool leader = warp == 0 && lane == 0;

    uint32_t gpage = 0;      // global index of the current task's first page
    uint32_t j = 0;          // sequence number of the next staged record
    int head = 0;            // ring row the next staged record starts at
    F T[W];
    long long wait_cycles = 0;  // tracing only

    for (uint32_t q = 0;; ++q) {
        mbar_wait<CHB_AB_COMPUTE_BACKOFF>(sm.tfull + q % kTaskQueue, (q / kTaskQueue) & 1);
        const QueuedTask cur = sm.queue[q % kTaskQueue];
        if (cur.ticket >= a.n_tasks) break;
        const int task = cur.task;
        const uint32_t total = cur.bytes;
        const uint32_t n_pages = (total + kPage - 1) / kPage;
        if (TRACE && leader) {
            a.trace[8 * (size_t) task] = global_ns();
            wait_cycles = 0;
        }

        int row0 = 0, w = 0;        // the panel whose tile is in registers (row0: global row, or local index)
        int lbase = 0, lrows = 0;   // ring rows of the local task being solved
        // waits for the loader to have staged the next record's rows; returns their ring position
        auto acquire_rows = [&](int nr) {
            const long long c0 = TRACE ? clock64() : 0;
            mbar_wait<CHB_AB_COMPUTE_BACKOFF>(sm.rfull + j % kRowBars, (j / kRowBars) & 1);
            if (TRACE) wait_cycles += clock64() - c0;
            if (head + nr > ring_rows) head = 0;  // same rule as the loader: a record's rows never wrap
            return head;
        };
        auto release_rows = [&](int nr, bool written = false) {
            // ring rows written with generic stores (local tasks only) are overwritten by TMA once released
#ifdef CHB_AB_RELEASE_FENCE_ALL
            written = true;
#endif
#ifdef CHB_AB_NO_RELEASE_FENCE
            written = false;
#endif
            if (written) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.rempty + j % kRowBars);
            j += 1;
            head += nr;
        };
        auto ring_at = [&](int base, int i) -> F & { return ring[(size_t) (base + i) * kp]; };
        // reference-order substitution: packed triangle, IEEE division (exact_order schedules)
        auto tri_exact = [&](const F *tv) {
            if (!EXACT) return;
#pragma unroll 1
            for (int step = 0; step < w; ++step) {
                const int s = UPPER ? w - 1 - step : step;
                const F d = *tv++;
                const F x = get_row<F>(T, s) / d;
                put_row<F>(T, s, x);
#pragma unroll
                for (int t = 0; t < W; ++t) {
                    if (UPPER ? t < s : (t > s && t < w)) {
                        const F l = *tv++;
                        T[t] = fnma(l, x, T[t]);
                    }
                }
            }
        };
        auto tri_fast = [&](const F *m) { tri_dense<F, UPPER>(T, m, w == 1 ? 1 : (w + 3) & ~3); };

        for (uint32_t pg = 0; pg < n_pages; ++pg) {
            const uint32_t g = gpage + pg, stage = g % kStages;
            mbar_wait<CHB_AB_COMPUTE_BACKOFF>(sm.full + stage, (g / kStages) & 1);
            const unsigned char *p = sm.slab + (size_t) stage * kPage;
            const unsigned char *pend = p + min((uint32_t) kPage, total - pg * kPage);
            while (p < pend) {
                const uint4 h = *reinterpret_cast<const uint4 *>(p);
                const uint32_t type = h.x & 0x3f;
                const unsigned char *pay = p + 16;
                p += (h.x >> 8) * 16;
                if (type == REC_BLOCK_D4) {
                    const int base = acquire_rows(__popc(h.z & 0xffff));
                    run_block<F, UPPER>(T, h.z, pay, ring, kp, base);
                    release_rows(__popc(h.z & 0xffff));
                } else if (type == REC_LOCAL_ROWS) {
                    // the rows of a local task as sparse rows, substituted in place in the ring (each lane its own column)
                    F *lrow = ring + (size_t) lbase * kp;
                    const unsigned char *q = pay;
                    for (int i = 0; i < (int) h.y; ++i) {
                        const uint32_t meta = *reinterpret_cast<const uint32_t *>(q);
                        const F pivot = *reinterpret_cast<const F *>(q + 8);
                        const int n4 = (int) (meta >> 8 & 0xff);
                        const unsigned char *idx = q + 16;
                        const F *val = reinterpret_cast<const F *>(q + 16 + ((4 * n4 + 15) & ~15));
                        F *dst = lrow + (size_t) (meta & 0xff) * kp;
                        F acc = (meta >> 16 & 1) ? F(0) : *dst;  // bit 16: a contribution row, accumulated from zero
                        for (int g = 0; g < n4; ++g) {
                            const uint32_t i4 = *reinterpret_cast<const uint32_t *>(idx + 4 * g);
                            const Vec4<F> vv = Vec4<F>::load(val + 4 * g);
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc = fnma(vv.x[j], lrow[(size_t) (i4 >> (8 * j) & 0xff) * kp], acc);
                        }
                        acc = EXACT ? acc / pivot : acc * pivot;
                        if (active) *dst = acc;
                        q = reinterpret_cast<const unsigned char *>(val + 4 * n4);
                    }
                } else if (!UPPER && type == REC_ADD_ROWS) {
                    // partial sums formed by the local tasks below: tile[t] += row, targets of tmask in ascending order
                    const int nt = (int) (h.z & 0xffff);
                    const uint32_t tmask = h.z >> 16;
                    const F *row = &ring_at(acquire_rows(nt), 0);
#pragma unroll
                    for (int t = 0; t < W; ++t) {
                        if (tmask >> t & 1) {
                            T[t] += *row;
                            row += kp;
                        }
                    }
                    release_rows(nt);
                } else if (type == REC_BEGIN_RHS) {
                    row0 = (int) h.y;
                    w = (int) h.z;
                    const int base = acquire_rows(w);
#pragma unroll
                    for (int i = 0; i < W; ++i) T[i] = i < w ? ring_at(base, i) : F(0);
                    release_rows(w);
                } else if (type == REC_END_DENSE || type == REC_END_FINAL || type == REC_END_STORE) {
                    // (triangle,) then the tile goes to the residual vector or to the work vector
                    // (and, backward, through P' to x): one copy of the store code for all of them
                    const int *perm = reinterpret_cast<const int *>(pay);
                    const bool to_residual = type == REC_END_STORE && h.w == 1;
                    if (type != REC_END_STORE) {
                        const F *m = reinterpret_cast<const F *>(pay + (UPPER ? 4 * W : 0));
                        if (EXACT) tri_exact(m); else tri_fast(m);
                    }
                    if (active) {
                        F *out = (to_residual ? a.rbuf : a.tmp) + (size_t) row0 * kp + col;
                        const bool scatter = UPPER && !to_residual;
#pragma unroll
                        for (int i = 0; i < W; ++i) {
                            if (i < w) {
                                __stcg(out + (size_t) i * kp, T[i]);
                                if (scatter) a.dst[(size_t) perm[i] * a.dst_ld + col] = T[i];
                            }
                        }
                    }
                } else if (type == REC_LOCAL_LOAD) {
                    lrows = (int) (h.z & 0xffff) + (int) (h.z >> 16);
                    lbase = acquire_rows(lrows);   // held until LOCAL_STORE
                } else if (type == REC_LOCAL_STORE) {
                    const int lo = (int) h.y, n_int = (int) (h.z & 0xffff), n_out = (int) (h.z >> 16);
                    const int *perm = reinterpret_cast<const int *>(pay);
                    if (active) {
                        for (int i = 0; i < n_int; ++i) {
                            const F v = ring_at(lbase, i);
                            __stcg(a.tmp + (size_t) (lo + i) * kp + col, v);
                            if (UPPER) a.dst[(size_t) perm[i] * a.dst_ld + col] = v;
                        }
                        // forward: the task's contributions to its ancestors' rows (local rows n_int ...) -> contribution buffer
                        if (!UPPER) {
                            F *out = a.cbuf + (size_t) h.w * kp + col;
                            for (int e = 0; e < n_out; ++e) __stcg(out + (size_t) e * kp, ring_at(lbase, n_int + e));
                        }
                    }
                    release_rows(lrows, true);
                } else if (type == REC_BEGIN_ZERO) {
                    row0 = (int) h.y;
                    w = (int) h.z;
#pragma unroll
                    for (int i = 0; i < W; ++i) T[i] = F(0);
                } else if (type == REC_BEGIN_SLOT || type == REC_ADD_SLOT) {
                    if (type == REC_BEGIN_SLOT) {
                        row0 = (int) h.y;
                        w = (int) (h.z & 0xff);
                    }
                    const int base = acquire_rows(W);
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        const F val = ring_at(base, i);
                        T[i] = type == REC_BEGIN_SLOT ? val : T[i] + val;
                    }
                    release_rows(W);
                } else if (type == REC_END_SLOT) {
                    F *slot = a.scratch + (size_t) (h.z >> 8) * W * kp;
                    if (active) {
#pragma unroll
                        for (int i = 0; i < W; ++i) __stcg(slot + (size_t) i * kp + col, T[i]);
                    }
                }
                // REC_NOP: nothing
            }
            __syncwarp();  // every lane is done reading this stage
            if (lane == 0) mbar_arrive(sm.empty + stage);
        }
        gpage += n_pages;

        // publish: every compute warp's stores of this task are done before the flag is released
        if (n_compute > 1) named_barrier(1, n_compute * 32); else __syncwarp();
        if (leader) {
            if (TRACE) {
                a.trace[8 * (size_t) task + 2] = global_ns();
                a.trace[8 * (size_t) task + 3] = (unsigned long long) wait_cycles;
            }
            st_release_gpu(a.done + task, a.epoch);
            mbar_arrive(sm.tempty + q % kTaskQueue);
        }
    }
}

template <typename F, bool UPPER, bool EXACT, bool TRACE>
__device__ __forceinline__ void sweep_worker(const SweepArgs<F> &a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_compute = (blockDim.x >> 5) - 1;  // the last warp is the loader
    const int kp = (a.k + 3) & ~3;
    const Smem<F> sm(smem_raw, kp, a.ring_rows);
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) *a.ticket_next = 0;
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(sm.full + s, 1);
            mbar_init(sm.empty + s, n_compute + 1);
        }
#pragma unroll
        for (int s = 0; s < kRowBars; ++s) {
            mbar_init(sm.rfull + s, 1);
            mbar_init(sm.rempty + s, n_compute);
        }
#pragma unroll
        for (int s = 0; s < kTaskQueue; ++s) {
            mbar_init(sm.tfull + s, 1);
            mbar_init(sm.tempty + s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == n_compute) loader_warp<F, UPPER>(a, sm, lane);
    else compute_warp<F, UPPER, EXACT, TRACE>(a, sm, warp, lane, n_compute);
}

}  // namespace sweep

// _exact: kernels for CHB_EXACT_ORDER schedules (packed triangles, IEEE division); the fast kernels do not
// carry that code (instruction footprint).  _trace: the CHB_TRACE=1 variants with the per-task timestamps.
// CTAs per SM the fast kernels are compiled for (register cap): measured on cfg4 with the local tasks gone to their own
// kernel (profiles/r02x_ctas_ab.log) 4 -> 5 -> 6 CTAs per SM = 10.48 -> 9.50 -> 9.43 ms; 64 registers is the last
// spill-free point in f32 (7 CTAs: 24 B of stack); f64: 4 CTAs (96 registers, cfg5 6.19 -> 5.68 ms), 5 spill (6.8 ms).
#ifndef CHB_SWEEP_MIN_CTAS_F32
#define CHB_SWEEP_MIN_CTAS_F32 6
#endif
#ifndef CHB_SWEEP_MIN_CTAS_F64
#define CHB_SWEEP_MIN_CTAS_F64 4
#endif
#define CHB_SWEEP(NAME, T, UP, EX, MINB)                                                           \
    extern "C" __global__ void __launch_bounds__(160, MINB) NAME(const chb::SweepArgs<T> a) {       \
        sweep::sweep_worker<T, UP, EX, false>(a);                                                  \
    }                                                                                              \
    extern "C" __global__ void __launch_bounds__(160, MINB) NAME##_trace(const chb::SweepArgs<T> a) { \
        sweep::sweep_worker<T, UP, EX, true>(a);                                                   \
    }

CHB_SWEEP(chb_sweep_f32_lower, float, false, false, CHB_SWEEP_MIN_CTAS_F32)
CHB_SWEEP(chb_sweep_f32_upper, float, true, false, CHB_SWEEP_MIN_CTAS_F32)
CHB_SWEEP(chb_sweep_f64_lower, double, false, false, CHB_SWEEP_MIN_CTAS_F64)
CHB_SWEEP(chb_sweep_f64_upper, double, true, false, CHB_SWEEP_MIN_CTAS_F64)
CHB_SWEEP(chb_sweep_f32_lower_exact, float, false, true, 2)
CHB_SWEEP(chb_sweep_f32_upper_exact, float, true, true, 2)
CHB_SWEEP(chb_sweep_f64_lower_exact, double, false, true, 2)
CHB_SWEEP(chb_sweep_f64_upper_exact, double, true, true, 2)
