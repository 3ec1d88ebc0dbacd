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
//   * LOCAL tasks (complete subtrees of small panels, most of the rows): the subtree's right-hand
//     side rows and the external rows it reads are gathered into shared memory once (one batch of
//     independent loads), every record then works out of shared memory, and the solved rows are
//     written back once -- no per-row round trips to L2;
//   * dependencies: one epoch-valued ready flag per TASK (ld.acquire.gpu / st.release.gpu after a
//     named barrier of the compute warps); flags are never cleared (the epoch is the solve counter);
//   * the work vector `tmp` and the partial-tile scratch have their row stride padded to 4 elements
//     so every row is a whole number of 16-byte units (TMA / cp.async granularity);
//   * P is fused into the forward tile load, P' into the backward tile store.
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
constexpr int kTaskQueue = 4;    // task descriptors the loader may run ahead of the compute warps
constexpr unsigned kFullMask = 0xffffffffu;

enum : uint32_t { REC_NOP = 0, REC_BEGIN_RHS, REC_BEGIN_ZERO, REC_BEGIN_SLOT, REC_ADD_SLOT, REC_BLOCK_D4, REC_BLOCK_D1,
                  REC_BLOCK_TM, REC_END_FINAL, REC_END_SLOT, REC_END_DENSE, REC_LOCAL_LOAD, REC_LOCAL_STORE,
                  REC_LOCAL_FLAG = 0x40 };

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    unsigned spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
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

// Substitution with the panel's triangle given as a dense WP x WP matrix with reciprocal pivots
// (REC_END_DENSE): fully static, one row of WP values per pivot from shared memory.
template <typename F, bool UPPER, int WP>
__device__ __forceinline__ void tri_dense(F (&T)[W], const F *m) {
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
        T[s] = T[s] * row[s];
#pragma unroll
        for (int t = UPPER ? 0 : s + 1; t < (UPPER ? s : WP); ++t) T[t] = fnma(row[t], T[s], T[t]);
    }
}

__device__ __forceinline__ bool is_staged_block(uint32_t h0) {  // a block record outside a local task
    const uint32_t type = h0 & 0xff;
    return type >= REC_BLOCK_D4 && type <= REC_BLOCK_TM;
}
// index (bit position) of the r-th source of a mask in sweep order (ascending forward, descending backward)
__device__ __forceinline__ int nth_source(uint32_t mask, int r, bool upper) {
    if (upper) mask = __brev(mask) >> 16;  // 16-bit masks
    const int pos = __fns(mask, 0, r + 1);
    return upper ? 15 - pos : pos;
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// Spin loops trap instead of hanging the GPU if a dependency never arrives (a schedule or kernel bug).
#define SWEEP_SPIN_GUARD(counter)            \
    if (++(counter) > (1u << 27)) __trap();

// A task descriptor in the CTA's shared-memory task queue (loader -> compute warps).
struct QueuedTask {
    int ticket;   // >= n_tasks: no more work
    int task;     // id = index of the task's ready flag
    unsigned int bytes;
    unsigned int pad;
    unsigned long long offset;
    unsigned long long pad2;
};

// One block record: tile -= L_block * sources.  Source row r (sweep order) is row (base + r) mod ring_rows
// of the row ring (STAGED), or row idx[r] of the local area (LOCAL; idx = 16 bytes in front of the values).
// `rows` points at this lane's column of row 0; rows are kp elements apart.
template <typename F, bool LOCAL>
__device__ __forceinline__ void run_block(F (&T)[W], uint32_t type, uint32_t masks, const unsigned char *pay, const F *rows,
                                          int kp, int base, int ring_rows) {
    const uint32_t tmask = masks >> 16, smask = masks & 0xffff;
    const unsigned char *idx = pay;
    const F *val = reinterpret_cast<const F *>(pay + (LOCAL ? 16 : 0));
    const int ns = __popc(smask);
    auto source = [&](int r) -> F {
        if (LOCAL) return rows[(int) idx[r] * kp];
        int q = base + r;
        if (q >= ring_rows) q -= ring_rows;
        return rows[q * kp];
    };
    if (type == REC_BLOCK_D4) {
        const int stride = (32 - __clz(tmask) + 3) & ~3;
        if (stride == 16) {
#pragma unroll 2
            for (int r = 0; r < ns; ++r) {
                const F x = source(r);
                const F *vr = val + r * 16;
#pragma unroll
                for (int t4 = 0; t4 < 4; ++t4) {
                    const Vec4<F> vv = Vec4<F>::load(vr + t4 * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) T[t4 * 4 + j] = fnma(vv.x[j], x, T[t4 * 4 + j]);
                }
            }
        } else {
            for (int r = 0; r < ns; ++r) {
                const F x = source(r);
                const F *vr = val + r * stride;
#pragma unroll
                for (int t4 = 0; t4 < 3; ++t4) {
                    if (t4 * 4 < stride) {
                        const Vec4<F> vv = Vec4<F>::load(vr + t4 * 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) T[t4 * 4 + j] = fnma(vv.x[j], x, T[t4 * 4 + j]);
                    }
                }
            }
        }
    } else if (type == REC_BLOCK_D1) {
        const int nt = __popc(tmask);
        for (int r = 0; r < ns; ++r) {
            const F x = source(r);
            const F *vr = val + r * nt;
#pragma unroll
            for (int t = 0; t < 3; ++t)
                if (t < nt) T[t] = fnma(vr[t], x, T[t]);
        }
    } else {  // REC_BLOCK_TM: target-major, arbitrary target set
        uint32_t tm = tmask;
        while (tm) {
            const int t = __ffs(tm) - 1;
            tm &= tm - 1;
            F acc = get_row<F>(T, t);
            for (int r = 0; r < ns; ++r) acc = fnma(val[r], source(r), acc);
            val += ns;
            put_row<F>(T, t, acc);
        }
    }
}

// Records whose input rows the LOADER stages into the row ring (everything else is handled by the
// compute warps alone).  Returns the number of rows, 0 if the record is not staged.
template <bool UPPER>
__device__ __forceinline__ int staged_rows(const uint4 &h) {
    const uint32_t type = h.x & 0xff;  // local records carry REC_LOCAL_FLAG and never match
    if (type >= REC_BLOCK_D4 && type <= REC_BLOCK_TM) return __popc(h.z & 0xffff);
    if (type == REC_BEGIN_SLOT || type == REC_ADD_SLOT) return W;
    if (UPPER && type == REC_BEGIN_RHS) return (int) h.z;
    return 0;
}

// Shared memory of one CTA (dynamic):
//   [kStages][kPage]          slab ring (TMA destination)
//   [ring_rows][kp]           row ring: already-solved rows / partial tiles staged by the loader
//   [local_rows][kp]          local area: the rows of the local task being solved
//   full[kStages], empty[kStages]      slab ring barriers   (empty: compute warps + loader)
//   rfull[kRowBars], rempty[kRowBars]  row ring barriers    (rempty: compute warps)
//   tfull[kTaskQueue], tempty[kTaskQueue]
//   QueuedTask[kTaskQueue], int rows_of[kRowBars]
template <typename F>
struct Smem {
    unsigned char *slab;
    F *ring;
    F *area;
    uint64_t *full, *empty, *rfull, *rempty, *tfull, *tempty;
    QueuedTask *queue;
    int *rows_of;
    __device__ __forceinline__ Smem(unsigned char *raw, int kp, int ring_rows, int local_rows) {
        slab = raw;
        ring = reinterpret_cast<F *>(raw + (size_t) kStages * kPage);
        area = ring + (size_t) ring_rows * kp;
        full = reinterpret_cast<uint64_t *>(area + (size_t) local_rows * kp);
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
// polls the ready flag of its producer and copies its input rows into the row ring with 1-D TMA bulk
// copies completing on that record's mbarrier.  Runs ahead of the compute warps by up to the depth of
// the three rings, across task boundaries.
template <typename F, bool UPPER>
__device__ __forceinline__ void loader_warp(const SweepArgs<F> &a, const Smem<F> &sm, int lane, int n_compute) {
    const int k = a.k, kp = (k + 3) & ~3;
    const int ring_rows = a.ring_rows;
    const uint32_t row_bytes = (uint32_t) kp * (uint32_t) sizeof(F);
    const int epoch = a.epoch;
    uint32_t gpage = 0;       // global index of the current task's first page
    uint32_t issued = 0;      // pages handed to TMA so far (lane 0 only)
    uint32_t j_issue = 0;     // staged records issued so far == sequence number of the next one
    uint32_t rows_issued = 0, rows_released = 0, j_rel = 0;
    int t_next = 0;
    if (lane == 0) t_next = atomicAdd(a.ticket, 1);

    for (uint32_t q = 0;; ++q) {
        // ---- next task descriptor -> queue --------------------------------------------------------
        int ticket = 0, task = 0;
        uint32_t total = 0, off_lo = 0, off_hi = 0;
        if (lane == 0) {
            ticket = t_next;
            if (ticket < a.n_tasks) {
                const SweepTask td = a.tasks[ticket];
                task = (int) td.level, total = td.bytes, off_lo = (uint32_t) td.offset, off_hi = (uint32_t) (td.offset >> 32);
                t_next = atomicAdd(a.ticket, 1);  // in flight while this task is staged
            }
            if (q >= kTaskQueue) mbar_wait(sm.tempty + q % kTaskQueue, ((q / kTaskQueue) - 1) & 1);
            sm.queue[q % kTaskQueue] = QueuedTask{ticket, task, total, 0u, (unsigned long long) off_lo | (unsigned long long) off_hi << 32, 0ull};
            mbar_arrive(sm.tfull + q % kTaskQueue);
        }
        ticket = __shfl_sync(kFullMask, ticket, 0);
        if (ticket >= a.n_tasks) break;
        total = __shfl_sync(kFullMask, total, 0);
        off_lo = __shfl_sync(kFullMask, off_lo, 0);
        off_hi = __shfl_sync(kFullMask, off_hi, 0);
        const unsigned char *base = a.slab + ((unsigned long long) off_lo | (unsigned long long) off_hi << 32);
        const uint32_t n_pages = (total + kPage - 1) / kPage;

        for (uint32_t pg = 0; pg < n_pages; ++pg) {
            const uint32_t g = gpage + pg, stage = g % kStages;
            if (lane == 0) {
                // keep the page ring full: pages g .. g + kStages - 1 may be resident; only page g is waited for
                const uint32_t limit = min(gpage + n_pages, g + kStages);
                while (issued < limit) {
                    const uint32_t st = issued % kStages, fill = issued / kStages;
                    if (fill > 0) {
                        if (issued == g) mbar_wait(sm.empty + st, (fill - 1) & 1);
                        else if (!mbar_test(sm.empty + st, (fill - 1) & 1)) break;
                    }
                    const uint32_t pi = issued - gpage;
                    const uint32_t bytes = min((uint32_t) kPage, total - pi * kPage);
                    mbar_expect_tx(sm.full + st, bytes);
                    tma_load_1d(sm.slab + (size_t) st * kPage, base + (size_t) pi * kPage, bytes, sm.full + st);
                    issued += 1;
                }
            }
            __syncwarp();
            mbar_wait(sm.full + stage, (g / kStages) & 1);
            const unsigned char *p = sm.slab + (size_t) stage * kPage;
            const unsigned char *pend = p + min((uint32_t) kPage, total - pg * kPage);
            while (p < pend) {
                // scan: the next (up to 32) staged records of this page; lane i keeps record i
                int n = 0;
                uint32_t my_off = 0, my_code = 0;
                const unsigned char *qp = p;
                while (qp < pend && n < 32) {
                    const uint4 hq = *reinterpret_cast<const uint4 *>(qp);
                    if (staged_rows<UPPER>(hq) > 0) {
                        if (lane == n) {
                            my_off = (uint32_t) (qp - p);
                            my_code = (hq.x & 0xff) == REC_BEGIN_RHS ? 0u : hq.w;
                        }
                        n += 1;
                    }
                    qp += (hq.x >> 8) * 16;
                }
                if (n == 0) break;  // nothing left to stage in this page
                // poll: one flag per lane, the ready prefix goes now
                int n_go;
                {
                    unsigned spins = 0;
                    for (;;) {
                        bool ready = true;
                        if (lane < n && my_code != 0) ready = ld_acquire_gpu(a.done + (my_code - 1)) == epoch;
                        const unsigned pending = __ballot_sync(kFullMask, !ready);
                        n_go = pending ? __ffs(pending) - 1 : n;
                        if (n_go > 0) break;
                        SWEEP_SPIN_GUARD(spins)
                    }
                }
                const unsigned char *last = p;
                for (int i = 0; i < n_go; ++i) {
                    const unsigned char *rec = p + __shfl_sync(kFullMask, my_off, i);
                    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
                    const uint32_t type = h.x & 0xff;
                    const int nr = staged_rows<UPPER>(h);
                    // back-pressure: a free barrier pair and nr free rows in the ring
                    while (j_issue - j_rel >= (uint32_t) kRowBars || rows_issued + (uint32_t) nr - rows_released > (uint32_t) ring_rows) {
                        mbar_wait(sm.rempty + j_rel % kRowBars, (j_rel / kRowBars) & 1);
                        rows_released += (uint32_t) sm.rows_of[j_rel % kRowBars];
                        j_rel += 1;
                    }
                    __syncwarp();
                    uint64_t *bar = sm.rfull + j_issue % kRowBars;
                    if (lane == 0) {
                        sm.rows_of[j_issue % kRowBars] = nr;
                        mbar_expect_tx(bar, (uint32_t) nr * row_bytes);
                    }
                    if (lane < nr) {
                        const F *src;
                        if (type == REC_BEGIN_RHS) src = a.tmp + (size_t) ((int) h.y + lane) * kp;
                        else if (type == REC_BEGIN_SLOT || type == REC_ADD_SLOT) src = a.scratch + ((size_t) (h.z >> 8) * W + lane) * kp;
                        else src = a.tmp + (size_t) ((int) h.y + nth_source(h.z & 0xffff, lane, UPPER)) * kp;
                        const uint32_t slot = (rows_issued + (uint32_t) lane) % (uint32_t) ring_rows;
                        tma_load_1d(sm.ring + (size_t) slot * kp, src, row_bytes, bar);
                    }
                    __syncwarp();
                    j_issue += 1;
                    rows_issued += (uint32_t) nr;
                    last = rec + (h.x >> 8) * 16;
                }
                p = last;
            }
            __syncwarp();  // every lane is done reading this stage
            if (lane == 0) mbar_arrive(sm.empty + stage);
        }
        gpage += n_pages;
    }
}

// ---- the compute warps -----------------------------------------------------------------------------
template <typename F, bool UPPER>
__device__ __forceinline__ void compute_warp(const SweepArgs<F> &a, const Smem<F> &sm, int warp, int lane, int n_compute) {
    const int k = a.k, kp = (k + 3) & ~3;
    const int ring_rows = a.ring_rows;
    const int col = warp * 32 + lane;
    const bool active = col < k;
    const int kw = min(32, k - warp * 32);                        // columns owned by this warp
    const int chunks = ((kw + 3) & ~3) * (int) sizeof(F) / 16;    // 16-byte chunks of this warp's row segment
    const int lcol = active ? col : warp * 32;                    // a column this lane may always read
    const F *ring = sm.ring + lcol;
    F *area = sm.area + lcol;
    const int epoch = a.epoch;
    const bool leader = warp == 0 && lane == 0;

    uint32_t gpage = 0;      // global index of the current task's first page
    uint32_t j = 0;          // sequence number of the next staged record
    uint32_t rows_seen = 0;  // rows of all staged records so far (ring position of the next one)
    F T[W];
    long long wait_cycles = 0;  // tracing only

    for (uint32_t q = 0;; ++q) {
        mbar_wait(sm.tfull + q % kTaskQueue, (q / kTaskQueue) & 1);
        const QueuedTask cur = sm.queue[q % kTaskQueue];
        if (cur.ticket >= a.n_tasks) break;
        const int task = cur.task;
        const uint32_t total = cur.bytes;
        const uint32_t n_pages = (total + kPage - 1) / kPage;
        if (a.trace && leader) {
            a.trace[8 * (size_t) task] = global_ns();
            wait_cycles = 0;
        }

        int row0 = 0, w = 0;   // the panel whose tile is in registers (row0: global row, or local index)
        // waits for the loader to have staged the next record's rows; returns their ring position
        auto acquire_rows = [&]() {
            const long long c0 = a.trace ? clock64() : 0;
            mbar_wait(sm.rfull + j % kRowBars, (j / kRowBars) & 1);
            if (a.trace) wait_cycles += clock64() - c0;
            return (int) (rows_seen % (uint32_t) ring_rows);
        };
        auto release_rows = [&](int nr) {
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.rempty + j % kRowBars);
            j += 1;
            rows_seen += (uint32_t) nr;
        };
        auto ring_row = [&](int base, int i) -> F {
            int r = base + i;
            if (r >= ring_rows) r -= ring_rows;
            return ring[(size_t) r * kp];
        };
        // tile -> tmp rows (and, backward, -> x through P')
        auto store_tile = [&](const int *perm) {
#pragma unroll
            for (int i = 0; i < W; ++i) {
                if (i < w && active) {
                    __stcg(a.tmp + (size_t) (row0 + i) * kp + col, T[i]);
                    if (UPPER) a.dst[(size_t) perm[i] * k + col] = T[i];
                }
            }
        };
        // reference-order substitution: packed triangle, IEEE division (exact_order schedules)
        auto tri_exact = [&](const F *tv) {
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
        auto tri_fast = [&](const F *m) {
            if (w == 1) tri_dense<F, UPPER, 1>(T, m);
            else if (w <= 4) tri_dense<F, UPPER, 4>(T, m);
            else if (w <= 8) tri_dense<F, UPPER, 8>(T, m);
            else if (w <= 12) tri_dense<F, UPPER, 12>(T, m);
            else tri_dense<F, UPPER, 16>(T, m);
        };

        for (uint32_t pg = 0; pg < n_pages; ++pg) {
            const uint32_t g = gpage + pg, stage = g % kStages;
            mbar_wait(sm.full + stage, (g / kStages) & 1);
            const unsigned char *p = sm.slab + (size_t) stage * kPage;
            const unsigned char *pend = p + min((uint32_t) kPage, total - pg * kPage);
            while (p < pend) {
                const uint4 h = *reinterpret_cast<const uint4 *>(p);
                const uint32_t type = h.x & 0x3f;
                const bool local = (h.x & REC_LOCAL_FLAG) != 0;
                const unsigned char *pay = p + 16;
                p += (h.x >> 8) * 16;
                if (type >= REC_BLOCK_D4 && type <= REC_BLOCK_TM) {
                    if (local) {
                        run_block<F, true>(T, type, h.z, pay, area, kp, 0, 0);
                    } else {
                        const int base = acquire_rows();
                        run_block<F, false>(T, type, h.z, pay, ring, kp, base, ring_rows);
                        release_rows(__popc(h.z & 0xffff));
                    }
                } else if (type == REC_BEGIN_RHS) {
                    row0 = (int) h.y;
                    w = (int) h.z;
                    if (local) {
#pragma unroll
                        for (int i = 0; i < W; ++i) T[i] = i < w ? area[(size_t) (row0 + i) * kp] : F(0);
                    } else if (UPPER) {
                        const int base = acquire_rows();
#pragma unroll
                        for (int i = 0; i < W; ++i) T[i] = i < w ? ring_row(base, i) : F(0);
                        release_rows(w);
                    } else {
                        const int *perm = reinterpret_cast<const int *>(pay);
#pragma unroll
                        for (int i = 0; i < W; ++i) T[i] = (i < w && active) ? __ldg(a.src + (size_t) perm[i] * k + col) : F(0);
                    }
                } else if (type == REC_END_DENSE || type == REC_END_FINAL) {
                    const int *perm = reinterpret_cast<const int *>(pay);
                    const F *m = reinterpret_cast<const F *>(pay + (UPPER && !local ? 4 * W : 0));
                    if (type == REC_END_DENSE) tri_fast(m); else tri_exact(m);
                    if (local) {
                        if (active) {
#pragma unroll
                            for (int i = 0; i < W; ++i)
                                if (i < w) area[(size_t) (row0 + i) * kp] = T[i];
                        }
                    } else {
                        store_tile(perm);
                    }
                } else if (type == REC_LOCAL_LOAD) {
                    const int lo = (int) h.y, n_int = (int) (h.z & 0xffff), n_ext = (int) (h.z >> 16), n_wait = (int) h.w;
                    const int *list = reinterpret_cast<const int *>(pay);
                    if (!UPPER) {
                        // area[i] = b[perm[i]]: independent loads, 8 in flight per lane
                        if (active) {
                            for (int i0 = 0; i0 < n_int; i0 += 8) {
                                F v[8];
#pragma unroll
                                for (int jj = 0; jj < 8; ++jj)
                                    v[jj] = i0 + jj < n_int ? __ldg(a.src + (size_t) list[i0 + jj] * k + col) : F(0);
#pragma unroll
                                for (int jj = 0; jj < 8; ++jj)
                                    if (i0 + jj < n_int) area[(size_t) (i0 + jj) * kp] = v[jj];
                            }
                        }
                    } else {
                        const int *waits = list + ((n_ext + 3) & ~3);
                        const long long c0 = a.trace ? clock64() : 0;
                        for (int e = lane; e < n_wait; e += 32) {
                            const int *flag = a.done + (waits[e] - 1);
                            unsigned spins = 0;
                            while (ld_acquire_gpu(flag) != epoch) { SWEEP_SPIN_GUARD(spins) }
                        }
                        __syncwarp();
                        if (a.trace) wait_cycles += clock64() - c0;
                        const int n_rows = n_int + n_ext;
                        F *dst0 = sm.area + warp * 32;
                        const F *src0 = a.tmp + warp * 32;
                        for (int e = lane; e < n_rows * chunks; e += 32) {
                            const int r = e / chunks, c = e - r * chunks;
                            const int grow = r < n_int ? lo + r : list[r - n_int];
                            cp_async_16(reinterpret_cast<unsigned char *>(dst0 + (size_t) r * kp) + 16 * c,
                                        reinterpret_cast<const unsigned char *>(src0 + (size_t) grow * kp) + 16 * c);
                        }
                        cp_async_commit();
                        cp_async_wait<0>();
                        __syncwarp();
                    }
                } else if (type == REC_LOCAL_STORE) {
                    const int lo = (int) h.y, n_int = (int) h.z;
                    const int *perm = reinterpret_cast<const int *>(pay);
                    if (active) {
                        for (int i = 0; i < n_int; ++i) {
                            const F v = area[(size_t) i * kp];
                            __stcg(a.tmp + (size_t) (lo + i) * kp + col, v);
                            if (UPPER) a.dst[(size_t) perm[i] * k + col] = v;
                        }
                    }
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
                    const int base = acquire_rows();
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        const F val = ring_row(base, i);
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
            if (a.trace) {
                a.trace[8 * (size_t) task + 2] = global_ns();
                a.trace[8 * (size_t) task + 3] = (unsigned long long) wait_cycles;
            }
            st_release_gpu(a.done + task, epoch);
            mbar_arrive(sm.tempty + q % kTaskQueue);
        }
    }
}

template <typename F, bool UPPER>
__device__ __forceinline__ void sweep_worker(const SweepArgs<F> &a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_compute = (blockDim.x >> 5) - 1;  // the last warp is the loader
    const int kp = (a.k + 3) & ~3;
    const Smem<F> sm(smem_raw, kp, a.ring_rows, a.local_rows);
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
    if (warp == n_compute) loader_warp<F, UPPER>(a, sm, lane, n_compute);
    else compute_warp<F, UPPER>(a, sm, warp, lane, n_compute);
}

}  // namespace sweep

#define CHB_SWEEP(NAME, T, UP)                                                                 \
    extern "C" __global__ void __launch_bounds__(160) NAME(const chb::SweepArgs<T> a) {          \
        sweep::sweep_worker<T, UP>(a);                                                         \
    }

CHB_SWEEP(chb_sweep_f32_lower, float, false)
CHB_SWEEP(chb_sweep_f32_upper, float, true)
CHB_SWEEP(chb_sweep_f64_lower, double, false)
CHB_SWEEP(chb_sweep_f64_upper, double, true)
