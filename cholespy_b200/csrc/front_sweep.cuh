// Front sweep kernels (sm_100a): forward (L y = P b) and backward (L' x = y, x -> P' x) triangular solves over the
// supernodal task image built by front_schedule.cpp.  One kernel serves both sweeps (the record stream says what
// to do); variants differ in element type and in the number of RHS columns a lane owns.
//
// What they replace: solve_lower_* / solve_upper_* of the reference (kernels/kernels.cu:64-131, 133-202: one
// 128-thread CTA per ROW, one flag poll + two block barriers per non-zero, one lane per RHS column).
//
// Execution model
//   * persistent co-resident grid; a CTA = 7 (or 4) COMPUTE warps + 1 LOADER warp; tickets are drawn from one atomic
//     counter in a topological order of the task graph (forward progress: the oldest unfinished ticket always
//     runs with its dependencies done);
//   * the loader streams the task's record slab into a ring of pages with 1-D TMA bulk copies (factor bytes leave
//     HBM exactly once per sweep), finds the staged records (FR_LOADX), polls their producers' ready flags
//     (ld.acquire.gpu) and copies the rows they name -- solved rows, right-hand-side rows of the user's b through
//     the permutation, rows of a local task -- into a shared-memory ROW RING with TMA bulk copies that complete
//     on the record's mbarrier.  It runs ahead of the compute warps by the depth of the rings, across task
//     boundaries;
//   * DENSE steps (FR_MMA): a warp owns a tile of TR rows x all RHS columns (CPL columns per lane: TR x CPL
//     accumulators in registers), and adds A(TR x s) X(s x n_rhs): X rows are read from the row ring (one
//     conflict-free vector load per source row), the TR factor values of a source column are broadcast loads
//     from the page.  Per source row and lane: 1 + TR/4 shared loads for TR x CPL fused multiply-adds
//     (f32, 128 RHS: 5 loads per 64 FMAs), no index arithmetic, no per-block bookkeeping;
//   * the epilogue (FR_EPI) writes tiles to the work vector / the update-row buffer / the user's x through P',
//     or back into ring rows; FR_ASM subtracts the update rows addressed to the staged right-hand side
//     (8 independent row loads in flight per warp); diagonal blocks are applied as explicit inverses (more MMAs);
//   * LOCAL tasks (complete subtrees): rows staged once, substituted in place as sparse rows -- dependency level by
//     level, the rows of a level dealt to the compute warps, one warp per row over all RHS columns -- and written
//     back once;
//   * dependencies: one epoch-valued ready flag per task, st.release.gpu after a barrier of the compute warps;
//     flags are never cleared (the epoch is the solve counter);
//   * proxies: rows written with generic stores and later read or overwritten by TMA are ordered with
//     fence.proxy.async on both sides (after the acquire poll in the loader, before a staged record is released
//     by the compute warps).
#pragma once
#include <cstdint>

#include "kernel_args.h"

namespace front {

using chb::FrontArgs;
using chb::SweepTask;

constexpr int kStages = 3;       // pages in the CTA's slab ring
constexpr int kRowBars = 16;     // staged records in flight
constexpr int kTaskQueue = 4;    // task descriptors the loader may run ahead
constexpr unsigned kFullMask = 0xffffffffu;

enum : uint32_t { FR_NOP = 0, FR_LOADX, FR_MMA, FR_EPI, FR_ASM, FR_BAR, FR_REL, FR_LROWS, FR_LSTORE, FR_DIR, FR_LTOP };
enum : uint32_t { FE_RING_SUB = 1, FE_RING_SET = 2, FE_Y = 4, FE_W = 8, FE_DST = 16 };

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
// 16-byte asynchronous copy global -> shared (LDGSTS); completion is reported to an mbarrier by cp_async_arrive.
// Used for GATHERED rows: one warp-wide instruction moves a 512-byte row, where a bulk copy per row costs a TMA
// request each (measured: ~50 ns per row, the fixed cost of a local task's 112-row right-hand side).
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// The mbarrier's pending count is incremented now and decremented when all prior cp.async of this thread are done.
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Watchdog: a dependency that never arrives (a schedule or kernel bug) raises the error word and ends the wait
// instead of hanging the GPU or killing the context; the host reports it after the sweep (GpuSolver::check_error).
constexpr unsigned long long kWatchdogNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *error_word) {
    // No suspend-time hint: ptxas turns a hinted try_wait into try + NANOSLEEP(hint) loops, which quantises every wait
    // to the hint (measured with a 20 us hint: 35 % of all warp samples sat in that NANOSLEEP).  The plain form blocks
    // in hardware for a short, implementation-defined time and wakes up on the phase flip.
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
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
        if ((spins & 4095u) == 4095u) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kWatchdogNs || *(volatile int *) error_word != 0) {
                atomicExch(error_word, 1);
                return;
            }
        }
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- CPL-wide vectors of one lane's columns ---------------------------------------------------------
template <typename F, int N>
struct alignas(sizeof(F) * N > 16 ? 16 : sizeof(F) * N) Vec {
    F v[N];
};
// L2-coherent global accesses (rows written by other SMs in this launch or in earlier launches: never through L1)
template <int BYTES>
__device__ __forceinline__ void ldcg_bytes(void *dst, const void *src) {
    if constexpr (BYTES == 4) {
        *(uint32_t *) dst = __ldcg((const uint32_t *) src);
    } else if constexpr (BYTES == 8) {
        *(uint2 *) dst = __ldcg((const uint2 *) src);
    } else if constexpr (BYTES == 16) {
        *(uint4 *) dst = __ldcg((const uint4 *) src);
    } else {
        ((uint4 *) dst)[0] = __ldcg((const uint4 *) src);
        ((uint4 *) dst)[1] = __ldcg((const uint4 *) src + 1);
    }
}
template <int BYTES>
__device__ __forceinline__ void stcg_bytes(void *dst, const void *src) {
    if constexpr (BYTES == 4) {
        __stcg((uint32_t *) dst, *(const uint32_t *) src);
    } else if constexpr (BYTES == 8) {
        __stcg((uint2 *) dst, *(const uint2 *) src);
    } else if constexpr (BYTES == 16) {
        __stcg((uint4 *) dst, *(const uint4 *) src);
    } else {
        __stcg((uint4 *) dst, ((const uint4 *) src)[0]);
        __stcg((uint4 *) dst + 1, ((const uint4 *) src)[1]);
    }
}
template <typename F, int N> __device__ __forceinline__ Vec<F, N> ld_global(const F *p) {
    Vec<F, N> r;
    ldcg_bytes<(int) sizeof(F) * N>(&r, p);
    return r;
}
template <typename F, int N> __device__ __forceinline__ void st_global(F *p, const Vec<F, N> &v) {
    stcg_bytes<(int) sizeof(F) * N>(p, &v);
}
template <typename F, int N> __device__ __forceinline__ Vec<F, N> ld_shared(const F *p) {
    return *reinterpret_cast<const Vec<F, N> *>(p);
}
template <typename F, int N> __device__ __forceinline__ void st_shared(F *p, const Vec<F, N> &v) {
    *reinterpret_cast<Vec<F, N> *>(p) = v;
}

template <typename F> struct Tile { static constexpr int rows = 16; };
template <> struct Tile<double> { static constexpr int rows = 8; };

struct QueuedTask {
    int ticket;  // >= n_tasks: no more work
    int task;    // id = index of the task's ready flag
    unsigned int bytes;
    unsigned int pad;
};

// Shared memory of one CTA (dynamic):
//   [kStages][page]       slab ring (TMA destination)
//   [ring_rows + 1][kp]   row ring
//   full/empty[kStages], rfull/rempty[kRowBars], tfull/tempty[kTaskQueue], queue, rows_of, rows_base
template <typename F>
struct Smem {
    unsigned char *slab;
    F *ring;
    uint64_t *full, *empty, *rfull, *rempty, *tfull, *tempty;
    QueuedTask *queue;
    int *rows_of, *rows_base;
    __device__ __forceinline__ Smem(unsigned char *raw, int kp, int ring_rows, int page) {
        slab = raw;
        ring = reinterpret_cast<F *>(raw + (size_t) kStages * page);
        full = reinterpret_cast<uint64_t *>(ring + (size_t) (ring_rows + 1) * kp);
        empty = full + kStages;
        rfull = empty + kStages;
        rempty = rfull + kRowBars;
        tfull = rempty + kRowBars;
        tempty = tfull + kTaskQueue;
        queue = reinterpret_cast<QueuedTask *>(tempty + kTaskQueue);
        rows_of = reinterpret_cast<int *>(queue + kTaskQueue);
        rows_base = rows_of + kRowBars;
    }
};

// ---- the loader warp -------------------------------------------------------------------------------
template <typename F, int NW, bool TRACE>
__device__ __forceinline__ void loader_warp(const FrontArgs<F> &a, const Smem<F> &sm, int lane) {
    const int k = a.k, kp = (k + 3) & ~3;
    const int ring_rows = a.ring_rows;
    const uint32_t page = (uint32_t) a.page_bytes;
    const uint32_t row_bytes = (uint32_t) kp * (uint32_t) sizeof(F);
    const uint32_t src_row_bytes = (uint32_t) k * (uint32_t) sizeof(F);
    const bool src_tma = a.src_tma != 0;
    const int epoch = a.epoch;
    uint32_t issued = 0;      // pages handed to TMA so far
    uint32_t j_issue = 0;     // staged records issued so far
    uint32_t j_rel = 0;       // staged records known to be released by the compute warps
    int rows_in_flight = 0;
    int head = 0;
    uint32_t q = 0;           // tasks queued so far
    int seen0 = -1, seen1 = -1;  // flags known to be set (this epoch)

    // Three tickets are in the loader's hands: the current task, the next one (descriptor known: its slab is
    // prefetched into L2 and its first pages into the ring while the current task runs) and the one after (ticket
    // drawn, descriptor load in flight).  Ticket and descriptor round trips are issued one task ahead of their use,
    // so neither is exposed: per-task loader latency is what bounds the throughput of small tasks.
    int t_cur = 0, t_next = 0, t_nn = 0;
    SweepTask d_cur{0ull, 0u, 0u}, d_next{0ull, 0u, 0u};
    uint32_t first_cur = 0, first_next = 0;
    bool next_known = false;  // d_next holds the descriptor of t_next (warp-uniform copy)
    SweepTask s_next{0ull, 0u, 0u};  // lane 0: descriptor load of t_next, possibly still in flight
    auto draw = [&]() {  // lane 0 only; the value is consumed much later
        return lane == 0 ? atomicAdd(a.ticket, 1) : 0;
    };
    auto load_desc = [&](int t) {
        SweepTask d{0ull, 0u, 0u};
        if (lane == 0 && t < a.n_tasks) d = a.tasks[t];
        return d;
    };
    auto broadcast = [&](const SweepTask &d) {
        SweepTask r;
        r.offset = __shfl_sync(kFullMask, d.offset, 0);
        r.bytes = __shfl_sync(kFullMask, d.bytes, 0);
        r.level = __shfl_sync(kFullMask, d.level, 0);
        return r;
    };
    auto prefetch_l2 = [&](const SweepTask &d) {  // the whole slab of a task (at most 1 MB), 32 KB per lane
        for (uint32_t off = (uint32_t) lane * 32768u; off < d.bytes && off < (1u << 20); off += 32u * 32768u) {
            const uint32_t bytes = min(32768u, d.bytes - off);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.slab + d.offset + off), "r"(bytes) : "memory");
        }
    };
    auto pages_of = [&](const SweepTask &d) { return (d.bytes + page - 1) / page; };
    // the next task's descriptor: made warp-uniform at its first use
    auto know_next = [&]() {
        if (next_known) return;
        d_next = broadcast(s_next);
        next_known = true;
        if (t_next < a.n_tasks && (TRACE && (a.debug & 64))) prefetch_l2(d_next);
    };
    // hands pages to TMA in global order while the ring has room; blocks only for page `must` (the one needed now)
    auto feed_pages = [&](uint32_t must) {
        for (;;) {
            if (issued >= must + kStages) break;
            // which task owns page `issued`?
            const SweepTask *d;
            uint32_t first;
            if (issued < first_cur + pages_of(d_cur)) {
                d = &d_cur, first = first_cur;
            } else {
                if (t_next >= a.n_tasks) break;
                know_next();
                if (issued >= first_next + pages_of(d_next)) break;
                d = &d_next, first = first_next;
            }
            const uint32_t st = issued % kStages, fill = issued / kStages;
            if (fill > 0) {
                if (issued <= must) {
                    mbar_wait(sm.empty + st, (fill - 1) & 1, a.error);
                } else {
                    const int free_now = lane == 0 ? (int) mbar_test(sm.empty + st, (fill - 1) & 1) : 0;
                    if (!__shfl_sync(kFullMask, free_now, 0)) break;
                }
            }
            const uint32_t pi = issued - first;
            const uint32_t bytes = min(page, d->bytes - pi * page);
            if (lane == 0) {
                mbar_expect_tx(sm.full + st, bytes);
                tma_load_1d(sm.slab + (size_t) st * page, a.slab + d->offset + (size_t) pi * page, bytes, sm.full + st);
            }
            issued += 1;
        }
    };

    t_cur = __shfl_sync(kFullMask, draw(), 0);
    d_cur = broadcast(load_desc(t_cur));
    if (t_cur < a.n_tasks && (TRACE && (a.debug & 64))) prefetch_l2(d_cur);
    t_next = __shfl_sync(kFullMask, draw(), 0);
    first_cur = 0;
    for (;;) {
        if (t_cur >= a.n_tasks) {
            if (lane == 0) {
                if (q >= kTaskQueue) mbar_wait(sm.tempty + q % kTaskQueue, ((q / kTaskQueue) - 1) & 1, a.error);
                sm.queue[q % kTaskQueue] = QueuedTask{a.n_tasks, 0, 0u, 0u};
                mbar_arrive(sm.tfull + q % kTaskQueue);
            }
            break;
        }
        const uint32_t n_pages = pages_of(d_cur);
        first_next = first_cur + n_pages;
        // in flight from here on, consumed later: the next task's descriptor and the ticket after it
        s_next = load_desc(t_next);
        next_known = false;
        t_nn = (TRACE && (a.debug & 256)) ? draw() : 0;
        if (lane == 0) {
            if (q >= kTaskQueue) mbar_wait(sm.tempty + q % kTaskQueue, ((q / kTaskQueue) - 1) & 1, a.error);
            sm.queue[q % kTaskQueue] = QueuedTask{t_cur, (int) d_cur.level, d_cur.bytes, 0u};
            mbar_arrive(sm.tfull + q % kTaskQueue);
        }
        q += 1;
        bool traced_first = false;
        const int tsel = TRACE ? (a.debug >> 10) & 3 : 0;  // which three loader events go to trace words 5..7
        auto stamp = [&](int sel, int word) {
            if (TRACE && a.trace && lane == 0 && tsel == sel) a.trace[8 * (size_t) d_cur.level + word] = global_ns();
        };
        stamp(0, 5);
        stamp(1, 5);
        for (uint32_t pg = 0; pg < n_pages; ++pg) {
            const uint32_t g = first_cur + pg, stage = g % kStages;
            feed_pages(g);
            __syncwarp();
            if (pg == 0) stamp(1, 6);
            mbar_wait(sm.full + stage, (g / kStages) & 1, a.error);
            if (pg == 0) stamp(1, 7);
            const unsigned char *p0 = sm.slab + (size_t) stage * page;
            const unsigned char *pend = p0 + min(page, d_cur.bytes - pg * page);
            // the page's directory (its first record) lists the staged records; more than six: walk the page
            const uint4 dir = *reinterpret_cast<const uint4 *>(p0);
            const uint32_t n_listed = dir.x >> 24;
            const unsigned char *p = p0 + 16;
            for (uint32_t li = 0;; ++li) {
                const unsigned char *rec;
                if (n_listed <= 6) {
                    if (li >= n_listed) break;
                    const uint32_t w = li < 2 ? dir.y : li < 4 ? dir.z : dir.w;
                    rec = p0 + 16u * ((li & 1u) ? w >> 16 : w & 0xffffu);
                } else {
                    rec = nullptr;
                    while (p < pend) {
                        const uint32_t h0 = *reinterpret_cast<const uint32_t *>(p);
                        const unsigned char *at = p;
                        p += (h0 >> 8 & 0xffffu) * 16;
                        if ((h0 & 0xff) == FR_LOADX) {
                            rec = at;
                            break;
                        }
                    }
                    if (!rec) break;
                }
                const uint4 h = *reinterpret_cast<const uint4 *>(rec);
                const int n_contig = (int) (h.y & 1023u), n_gather = (int) (h.y >> 10 & 1023u), n_extra = (int) (h.y >> 20 & 1023u);
                const int cbuf = (int) (h.w & 15u), gkind = (int) (h.w >> 4 & 15u), n_wait = (int) (h.w >> 8);
                const int *waits = reinterpret_cast<const int *>(rec + 16);
                const int *glist = waits + ((n_wait + 3) & ~3);
                // producers.  The last two flags seen set are remembered: consecutive chunks of one producer cost one poll.
                bool polled = false;
                if (!(n_wait == 1 && (waits[0] == seen0 || waits[0] == seen1)) && n_wait > 0) {
                    polled = true;
                    for (int e = lane; e < n_wait; e += 32) {
                        const int *flag = a.done + waits[e];
                        unsigned spins = 0;
                        unsigned long long t0 = 0;
                        while (ld_acquire_gpu(flag) != epoch) {
                            __nanosleep(100);  // thousands of loaders poll: keep their acquire loads (each invalidates L1) off the L2 queues
                            if ((++spins & 4095u) == 0) {
                                const unsigned long long now = global_ns();
                                if (t0 == 0) t0 = now;
                                if (now - t0 > kWatchdogNs || *(volatile int *) a.error != 0) {
                                    atomicExch(a.error, 2);
                                    break;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    seen1 = seen0;
                    seen0 = waits[0];
                }
                // the producers' generic-proxy stores are read through the async proxy (TMA) below
                if (polled) fence_proxy_async();
                // ring space: a record's rows never wrap around the ring end; back-pressure on rows and barrier pairs
                if (!traced_first) stamp(2, 5);
                const int nr = n_contig + n_gather + n_extra;
                int skipped = 0;
                if (head + nr > ring_rows) skipped = ring_rows - head, head = 0;
                while (j_issue - j_rel >= (uint32_t) kRowBars || (rows_in_flight > 0 && rows_in_flight + nr + skipped > ring_rows)) {
                    mbar_wait(sm.rempty + j_rel % kRowBars, (j_rel / kRowBars) & 1, a.error);
                    rows_in_flight -= sm.rows_of[j_rel % kRowBars];
                    j_rel += 1;
                }
                __syncwarp();
                if (!traced_first) stamp(2, 6);
                const uint32_t slot = j_issue % kRowBars;
                uint64_t *bar = sm.rfull + slot;
                if (lane == 0) {
                    sm.rows_of[slot] = nr + skipped;
                    sm.rows_base[slot] = head;
                }
                uint32_t tx = 0;
                if (n_contig > 0) {
                    const F *from = (cbuf == 1 ? a.tmp : a.wbuf) + (size_t) (int) h.z * kp;
                    if (lane == 0) tma_load_1d(sm.ring + (size_t) head * kp, from, (uint32_t) n_contig * row_bytes, bar);
                    tx += (uint32_t) n_contig * row_bytes;
                }
                F *gdst = sm.ring + (size_t) (head + n_contig) * kp;
                if ((gkind == 1 || (gkind == 2 && src_tma)) && !(TRACE && (a.debug & 128))) {
                    // gathered rows: one bulk copy per row, rows dealt to the lanes
                    const F *gsrc = gkind == 1 ? a.tmp : a.src;
                    const uint32_t bytes = gkind == 1 ? row_bytes : src_row_bytes;                                        // per row
                    const size_t stride = gkind == 1 ? (size_t) row_bytes : (size_t) a.src_ld * sizeof(F);  // between source rows
                    for (int i = lane; i < n_gather; i += 32)
                        tma_load_1d(gdst + (size_t) i * kp, reinterpret_cast<const unsigned char *>(gsrc) + (size_t) glist[i] * stride, bytes, bar);
                    tx += (uint32_t) n_gather * bytes;
                } else if (gkind == 1 || (gkind == 2 && src_tma)) {
                    // (experiment, CHB_FRONT_DEBUG bit 128) 16 bytes per lane and instruction (cp.async)
                    const F *gsrc = gkind == 1 ? a.tmp : a.src;
                    const size_t stride = gkind == 1 ? (size_t) row_bytes : (size_t) a.src_ld * sizeof(F);  // bytes between source rows
                    const int cpr = (int) ((gkind == 1 ? row_bytes : src_row_bytes) / 16u);  // 16-byte chunks per row
                    const unsigned char *gbase = reinterpret_cast<const unsigned char *>(gsrc) + 16 * lane;
                    unsigned char *dbase = reinterpret_cast<unsigned char *>(gdst) + 16 * lane;
                    for (int i0 = 0; i0 < n_gather; i0 += 32) {
                        const int mine = i0 + lane < n_gather ? glist[i0 + lane] : 0;  // one index per lane, handed around below
                        const int nrow = min(32, n_gather - i0);
#pragma unroll 4
                        for (int j = 0; j < nrow; ++j) {
                            const size_t from = (size_t) __shfl_sync(kFullMask, mine, j) * stride;
                            const size_t to = (size_t) (i0 + j) * row_bytes;
                            for (int c = lane; c < cpr; c += 32) cp_async_16(dbase + to + 16 * (c - lane), gbase + from + 16 * (c - lane));
                        }
                    }
                    cp_async_arrive(bar);
                } else if (gkind == 2) {
                    const int total = n_gather * k;
#pragma unroll 1
                    for (int e = lane; e < total; e += 32) {
                        const int i = e / k, c = e - i * k;
                        gdst[(size_t) i * kp + c] = __ldg(a.src + (size_t) glist[i] * a.src_ld + c);
                    }
                }
                __syncwarp();  // manual copies are complete before the arrival
                if (lane == 0) mbar_expect_tx(bar, tx);
                if (!traced_first) stamp(0, 6), stamp(2, 7);
                traced_first = true;
                j_issue += 1;
                rows_in_flight += nr + skipped;
                head += nr;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.empty + stage);
        }
        stamp(0, 7);
        // next task becomes current
        know_next();
        t_cur = t_next;
        d_cur = d_next;
        first_cur = first_next;
        t_next = __shfl_sync(kFullMask, (TRACE && (a.debug & 256)) ? t_nn : draw(), 0);
    }
}

// ---- the compute warps -----------------------------------------------------------------------------
template <typename F, int CPL, int NW, bool TRACE>
__device__ __forceinline__ void compute_warp(const FrontArgs<F> &a, const Smem<F> &sm, int warp, int lane) {
    constexpr int TR = Tile<F>::rows;
    constexpr int AV = 16 / (int) sizeof(F);     // factor values per 16-byte shared load
    using V = Vec<F, CPL>;
    const int k = a.k, kp = (k + 3) & ~3;
    const uint32_t page = (uint32_t) a.page_bytes;
    const int col = lane * CPL;
    const bool active = col < kp;
    const int lcol = active ? col : 0;           // a column this lane may always read
    const bool leader = warp == 0 && lane == 0;
    const bool dst_vec = a.dst_vec != 0;
    const int dbg = TRACE ? a.debug : 0;  // experiment switches exist in the tracing kernels only

    uint32_t gpage = 0;      // global index of the current task's first page
    uint32_t j0 = 0;         // sequence number of the task's first staged record
    // tracing only: ns spent by warp 0 per activity {0 queue wait, 1 page wait, 2 staged wait, 3 MMA, 4 EPI, 5 ASM, 6 BAR,
    // 7 REL, 8 LROWS, 9 LSTORE, 10 end barrier, 11 flag release, 12 other}
    unsigned long long prof[TRACE ? 13 : 1];
#pragma unroll
    for (int i = 0; i < (TRACE ? 13 : 1); ++i) prof[i] = 0;
    const bool profiling = TRACE && a.trace != nullptr && warp == 0;
    unsigned long long pt = profiling ? global_ns() : 0;
    auto lap = [&](int slot) {
        if (TRACE && profiling) {
            const unsigned long long now = global_ns();
#pragma unroll
            for (int i = 0; i < (TRACE ? 13 : 1); ++i)
                if (i == slot) prof[i] += now - pt;
            pt = now;
        }
    };

    // waits for staged record sj of the current task; returns its first ring row
    unsigned long long *t_first_ptr = nullptr, *t_wait_ptr = nullptr;
    uint32_t staged_j = 0xffffffffu;  // last staged record this warp waited for, and its first ring row
    int staged_base = 0;
    auto staged = [&](uint32_t sj) -> int {
        const uint32_t j = j0 + sj;
        if (j == staged_j) return staged_base;
        lap(12);
        if (TRACE && a.trace && warp == 0) {
            const unsigned long long c0 = global_ns();
            mbar_wait(sm.rfull + j % kRowBars, (j / kRowBars) & 1, a.error);
            const unsigned long long c1 = global_ns();
            *t_wait_ptr += c1 - c0;
            if (*t_first_ptr == 0) *t_first_ptr = c1;
        } else {
            mbar_wait(sm.rfull + j % kRowBars, (j / kRowBars) & 1, a.error);
        }
        staged_j = j;
        staged_base = sm.rows_base[j % kRowBars];
        lap(2);
        return staged_base;
    };
    auto store_dst = [&](int perm_row, const V &v) {
        F *d = a.dst + (size_t) perm_row * a.dst_ld + col;
        if (dst_vec) {
            *reinterpret_cast<V *>(d) = v;
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                if (col + c < k) d[c] = v.v[c];
        }
    };

    for (uint32_t q = 0;; ++q) {
        lap(12);
        mbar_wait(sm.tfull + q % kTaskQueue, (q / kTaskQueue) & 1, a.error);
        lap(0);
        const QueuedTask cur = sm.queue[q % kTaskQueue];
        if (cur.ticket >= a.n_tasks) break;
        const int task = cur.task;
        const uint32_t total = cur.bytes;
        const uint32_t n_pages = (total + page - 1) / page;
        unsigned long long t_first_data = 0, t_wait = 0;  // tracing only (warp 0)
        t_first_ptr = &t_first_data, t_wait_ptr = &t_wait;
        if (TRACE && a.trace && leader) {
            a.trace[8 * (size_t) task] = global_ns();
            a.trace[8 * (size_t) task + 1] = blockIdx.x;
        }
        uint32_t n_staged = 0;

        // record cursor over the task's pages
        uint32_t pg = 0, stage = 0;
        const unsigned char *p = nullptr, *pend = nullptr;
        bool page_open = false;
        auto next_record = [&](uint4 &h, const unsigned char *&pay) -> bool {
            while (p >= pend) {
                if (page_open) {
                    __syncwarp();  // every lane is done reading this stage
                    if (lane == 0) mbar_arrive(sm.empty + stage);
                    page_open = false;
                }
                if (pg == n_pages) return false;
                const uint32_t g = gpage + pg;
                stage = g % kStages;
                lap(12);
                mbar_wait(sm.full + stage, (g / kStages) & 1, a.error);
                lap(1);
                p = sm.slab + (size_t) stage * page;
                pend = p + min(page, total - pg * page);
                pg += 1;
                page_open = true;
            }
            h = *reinterpret_cast<const uint4 *>(p);
            pay = p + 16;
            p += (h.x >> 8 & 0xffffu) * 16;
            return true;
        };
        auto release = [&](uint32_t sj) {
            // every compute warp passes here, in staging order; the wait keeps a fast warp from arriving on a
            // barrier pair the loader has not recycled yet
            const uint32_t j = j0 + sj;
            mbar_wait(sm.rfull + j % kRowBars, (j / kRowBars) & 1, a.error);
            fence_proxy_async_smem();  // generic-proxy writes to these ring rows precede the TMA writes that reuse them
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.rempty + j % kRowBars);
        };

        uint4 h;
        const unsigned char *pay;
        while (next_record(h, pay)) {
            uint32_t type = h.x & 0xff;
            if (type == FR_MMA) {
                // One accumulation: MMA records (interleaved with staging / release / barrier records) up to the
                // epilogue.  The accumulators live only inside this block.
                F acc[TR][CPL];
                for (;;) {
                    if (type == FR_MMA) {
                        const int t = warp - (int) (h.y & 255u);
                        if (t >= 0 && t < (int) (h.y >> 8 & 255u) && !(dbg & 8)) {
                            const int n_src = (int) (h.y >> 16 & 255u);
                            const int base = staged(h.z & 0xffffu) + (int) (h.z >> 16);
                            if (h.y >> 24 & 1u) {
#pragma unroll
                                for (int i = 0; i < TR; ++i)
#pragma unroll
                                    for (int c = 0; c < CPL; ++c) acc[i][c] = F(0);
                            }
                            const F *A = reinterpret_cast<const F *>(pay) + (size_t) t * n_src * TR;
                            const F *X = sm.ring + (size_t) base * kp + lcol;
#pragma unroll 2
                            for (int s = 0; s < n_src; ++s) {
                                const V x = ld_shared<F, CPL>(X);
                                X += kp;
                                F av[TR];
#pragma unroll
                                for (int i = 0; i < TR; i += AV) {
                                    const Vec<F, AV> q4 = ld_shared<F, AV>(A + i);
#pragma unroll
                                    for (int j = 0; j < AV; ++j) av[i + j] = q4.v[j];
                                }
                                A += TR;
#pragma unroll
                                for (int i = 0; i < TR; ++i)
#pragma unroll
                                    for (int c = 0; c < CPL; ++c) acc[i][c] = fma(av[i], x.v[c], acc[i][c]);
                            }
                        }
                        lap(3);
                    } else if (type == FR_EPI) {
                        lap(12);
                        const int t = warp - (int) (h.y & 255u);
                        const uint32_t mode = h.y >> 16;
                        if (t >= 0 && t < (int) (h.y >> 8 & 255u) && !(dbg & 32)) {
                            const int r0 = t * TR;
                            const int nvalid = min(TR, (int) (h.z & 0xffffu) - r0);
                            const int *ip = reinterpret_cast<const int *>(pay);
                            F *ring_row = nullptr;
                            if (mode & (FE_RING_SUB | FE_RING_SET)) ring_row = sm.ring + (size_t) (staged(h.z >> 16) + ip[0] + r0) * kp + lcol;
                            const size_t g0 = (size_t) ((int) h.w + r0);
                            F *yrow = a.tmp + g0 * kp + lcol, *wrow = a.wbuf + g0 * kp + lcol;
#pragma unroll
                            for (int i = 0; i < TR; ++i) {
                                if (i < nvalid && active) {
                                    V v;
#pragma unroll
                                    for (int c = 0; c < CPL; ++c) v.v[c] = acc[i][c];
                                    if (mode & FE_RING_SUB) {
                                        const V o = ld_shared<F, CPL>(ring_row + (size_t) i * kp);
#pragma unroll
                                        for (int c = 0; c < CPL; ++c) v.v[c] = o.v[c] - v.v[c];
                                        st_shared<F, CPL>(ring_row + (size_t) i * kp, v);
                                    }
                                    if (mode & FE_RING_SET) st_shared<F, CPL>(ring_row + (size_t) i * kp, v);
                                    if (mode & FE_Y) st_global<F, CPL>(yrow + (size_t) i * kp, v);
                                    if (mode & FE_W) st_global<F, CPL>(wrow + (size_t) i * kp, v);
                                    if (mode & FE_DST) store_dst(ip[4 + r0 + i], v);
                                }
                            }
                        }
                        lap(4);
                        break;  // the epilogue ends the accumulation
                    } else if (type == FR_REL) {
                        lap(12);
                        release(h.y);
                        lap(7);
                    } else if (type == FR_LOADX) {
                        n_staged += 1;
                    } else if (type == FR_BAR) {
                        lap(12);
                        if (!(dbg & 4)) named_barrier(1, NW * 32);
                        lap(6);
                    }
                    if (!next_record(h, pay)) break;  // (malformed stream: an accumulation without epilogue)
                    type = h.x & 0xff;
                }
            } else if (type == FR_ASM) {
                lap(12);
                // ring[row] -= W[w_row] for this warp's entries, 8-16 row loads in flight
                const int *ip = reinterpret_cast<const int *>(pay);
                const int lo = ip[warp], hi = ip[warp + 1];
                if (lo < hi && !(dbg & 16)) {
                    F *rows = sm.ring + (size_t) (staged(h.y & 0xffffu) + (int) (h.y >> 16)) * kp + lcol;
                    const int2 *ent = reinterpret_cast<const int2 *>(ip + 12);
                    constexpr int kInFlight = 64 / ((int) sizeof(F) * CPL) >= 16 ? 16 : 8;  // row loads in flight (<= 64 registers)
                    for (int e = lo; e < hi; e += kInFlight) {
                        V vals[kInFlight];
                        int rr[kInFlight];
#pragma unroll
                        for (int j = 0; j < kInFlight; ++j) {
                            const int2 en = ent[min(e + j, hi - 1)];
                            rr[j] = en.x;
                            vals[j] = ld_global<F, CPL>(a.wbuf + (size_t) en.y * kp + lcol);
                        }
#pragma unroll
                        for (int j = 0; j < kInFlight; ++j) {
                            if (e + j < hi && active) {
                                F *dst = rows + (size_t) rr[j] * kp;
                                V o = ld_shared<F, CPL>(dst);
#pragma unroll
                                for (int c = 0; c < CPL; ++c) o.v[c] -= vals[j].v[c];
                                st_shared<F, CPL>(dst, o);
                            }
                        }
                    }
                }
                lap(5);
            } else if (type == FR_BAR) {
                lap(12);
                if (!(dbg & 4)) named_barrier(1, NW * 32);
                lap(6);
            } else if (type == FR_REL) {
                lap(12);
                release(h.y);
                lap(7);
            } else if (type == FR_LOADX) {
                n_staged += 1;
            } else if (type == FR_LROWS && !(dbg & 1)) {
                lap(12);
                // sparse rows of one dependency level of a local task: this warp substitutes its rows in place, over all
                // RHS columns (CPL per lane)
                const int *ip = reinterpret_cast<const int *>(pay);
                const unsigned char *qp = pay + ip[warp], *qend = pay + ip[warp + 1];
                if (qp < qend) {
                    F *lrow = sm.ring + (size_t) staged(h.z) * kp + lcol;
                    while (qp < qend) {
                        const uint32_t meta = *reinterpret_cast<const uint32_t *>(qp);
                        const F pivot = *reinterpret_cast<const F *>(qp + sizeof(F));
                        const int n4 = (int) (meta >> 8 & 0xff);
                        const unsigned char *idx = qp + 2 * sizeof(F);
                        const F *val = reinterpret_cast<const F *>(qp + ((2 * sizeof(F) + 4 * n4 + 15) & ~(size_t) 15));
                        F *dst = lrow + (size_t) (meta & 0xff) * kp;
                        V sv;
                        if (meta >> 16 & 1) {
#pragma unroll
                            for (int c = 0; c < CPL; ++c) sv.v[c] = F(0);
                        } else {
                            sv = ld_shared<F, CPL>(dst);
                        }
                        for (int g4 = 0; g4 < n4; ++g4) {
                            const uint32_t i4 = *reinterpret_cast<const uint32_t *>(idx + 4 * g4);
                            F vv[4];
#pragma unroll
                            for (int j = 0; j < 4; j += AV) {
                                const Vec<F, AV> v0 = ld_shared<F, AV>(val + 4 * g4 + j);
#pragma unroll
                                for (int jj = 0; jj < AV; ++jj) vv[j + jj] = v0.v[jj];
                            }
                            V xs[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) xs[j] = ld_shared<F, CPL>(lrow + (size_t) (i4 >> (8 * j) & 0xff) * kp);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int c = 0; c < CPL; ++c) sv.v[c] = fma(-vv[j], xs[j].v[c], sv.v[c]);
                        }
#pragma unroll
                        for (int c = 0; c < CPL; ++c) sv.v[c] *= pivot;
                        if (active) st_shared<F, CPL>(dst, sv);
                        qp = reinterpret_cast<const unsigned char *>(val + 4 * n4);
                    }
                }
                lap(8);
            } else if (type == FR_LTOP) {
                // the upper rows of a local task: x = inv(T) v as a dense product; this warp forms rows warp, warp + NW, ...
                // (at most 4) from the current values of all T rows, then -- after a barrier -- stores them
                const int T = (int) h.y;
                const unsigned char *idx = pay;
                const F *coef = reinterpret_cast<const F *>(pay + ((T + 15) & ~15)) + (size_t) warp * T * 4;
                F *lrow = sm.ring + (size_t) staged(h.z) * kp + lcol;
                V o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < CPL; ++c) o[i].v[c] = F(0);
                if (warp < T) {
#pragma unroll 2
                    for (int qq = 0; qq < T; ++qq) {
                        const V x = ld_shared<F, CPL>(lrow + (size_t) idx[qq] * kp);
                        F cf[4];
#pragma unroll
                        for (int j = 0; j < 4; j += AV) {
                            const Vec<F, AV> v0 = ld_shared<F, AV>(coef + 4 * qq + j);
#pragma unroll
                            for (int jj = 0; jj < AV; ++jj) cf[j + jj] = v0.v[jj];
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int c = 0; c < CPL; ++c) o[i].v[c] = fma(cf[i], x.v[c], o[i].v[c]);
                    }
                }
                named_barrier(1, NW * 32);
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (warp + i * NW < T) st_shared<F, CPL>(lrow + (size_t) idx[warp + i * NW] * kp, o[i]);
                }
            } else if (type == FR_LSTORE && !(dbg & 2)) {
                lap(12);
                const int n_int = (int) (h.y & 0xffffu), n_ext = (int) (h.y >> 16);
                const int *ip = reinterpret_cast<const int *>(pay);
                const F *rows = sm.ring + (size_t) staged(h.z) * kp + lcol;
                if (active) {
                    F *yrow = a.tmp + (size_t) (int) h.w * kp + col;
                    for (int i = warp; i < n_int; i += NW) {
                        const V v = ld_shared<F, CPL>(rows + (size_t) i * kp);
                        st_global<F, CPL>(yrow + (size_t) i * kp, v);
                        if (a.dst) store_dst(ip[4 + i], v);
                    }
                    F *wrow = a.wbuf + (size_t) ip[0] * kp + col;
                    for (int e = warp; e < n_ext; e += NW) st_global<F, CPL>(wrow + (size_t) e * kp, ld_shared<F, CPL>(rows + (size_t) (n_int + e) * kp));
                }
                lap(9);
            }
            // FR_NOP: nothing
        }
        gpage += n_pages;
        j0 += n_staged;

        // publish: every compute warp's stores of this task are done before the flag is released
        lap(12);
        named_barrier(1, NW * 32);
        lap(10);
        if (leader) {
            if (TRACE && a.trace) {
                a.trace[8 * (size_t) task + 2] = global_ns();
                a.trace[8 * (size_t) task + 3] = t_first_data;
                a.trace[8 * (size_t) task + 4] = t_wait;
            }
            st_release_gpu(a.done + task, a.epoch);
            mbar_arrive(sm.tempty + q % kTaskQueue);
        }
        lap(11);
    }
    if (TRACE && profiling && lane == 0) {
#pragma unroll
        for (int i = 0; i < (TRACE ? 13 : 1); ++i) a.trace[8 * (size_t) a.n_tasks + 16 * (size_t) blockIdx.x + i] = prof[i];
    }
}

template <typename F, int CPL, int NW, bool TRACE>
__device__ __forceinline__ void front_worker(const FrontArgs<F> &a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int kp = (a.k + 3) & ~3;
    const Smem<F> sm(smem_raw, kp, a.ring_rows, a.page_bytes);
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) *a.ticket_next = 0;
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(sm.full + s, 1);
            mbar_init(sm.empty + s, NW + 1);
        }
#pragma unroll
        for (int s = 0; s < kRowBars; ++s) {
            mbar_init(sm.rfull + s, 1);
            mbar_init(sm.rempty + s, NW);
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
    if (warp == NW) loader_warp<F, NW, TRACE>(a, sm, lane);
    else compute_warp<F, CPL, NW, TRACE>(a, sm, warp, lane);
}

}  // namespace front

// Two CTA shapes: 7 compute warps + loader (256 threads, 2 CTAs per SM) and 4 + loader (160 threads, 3 per SM);
// both leave 128 registers per thread (4 warps per scheduler).  The image is built for one of them (n_warps).
#define CHB_FRONT(NAME, T, CPL, NW, MINB)                                                                       \
    extern "C" __global__ void __launch_bounds__(32 * (NW + 1), MINB) NAME(const chb::FrontArgs<T> a) {          \
        front::front_worker<T, CPL, NW, false>(a);                                                              \
    }                                                                                                           \
    extern "C" __global__ void __launch_bounds__(32 * (NW + 1), MINB) NAME##_trace(const chb::FrontArgs<T> a) {  \
        front::front_worker<T, CPL, NW, true>(a);                                                               \
    }

CHB_FRONT(chb_front_f32_c1_w7, float, 1, 7, 2)
CHB_FRONT(chb_front_f32_c2_w7, float, 2, 7, 2)
CHB_FRONT(chb_front_f32_c4_w7, float, 4, 7, 2)
CHB_FRONT(chb_front_f64_c1_w7, double, 1, 7, 2)
CHB_FRONT(chb_front_f64_c2_w7, double, 2, 7, 2)
CHB_FRONT(chb_front_f64_c4_w7, double, 4, 7, 2)
CHB_FRONT(chb_front_f32_c1_w4, float, 1, 4, 3)
CHB_FRONT(chb_front_f32_c2_w4, float, 2, 4, 3)
CHB_FRONT(chb_front_f32_c4_w4, float, 4, 4, 3)
CHB_FRONT(chb_front_f64_c1_w4, double, 1, 4, 3)
CHB_FRONT(chb_front_f64_c2_w4, double, 2, 4, 3)
CHB_FRONT(chb_front_f64_c4_w4, double, 4, 4, 3)
