// Wide-level kernels (sm_100a): the panel tasks of one WIDE dependency level of a panel image executed by a plain
// grid, one launch per level, instead of by the persistent sweep kernel (sweep.cuh).
//
// Why: near the leaves of the elimination tree a dependency level holds thousands of small panel tasks (cfg4, forward:
// levels 1-10 hold 117 k tasks of 0.6-9 KB).  All of a level's tasks are independent and everything they read was
// written by earlier levels, so the level needs no ready flags at all -- but inside the persistent kernel each of these
// tasks pays the loader's serial costs per record (flag poll, row staging round trip, ring hand-over: ~3 us per block
// record per CTA, profiles/r02y_trace_cfg4_levels.txt: 2.4 ms for those ten levels).  Here a CTA of ceil(k / 32) warps
// owns a task, warp w the RHS columns [32 w, 32 w + 32) like the compute warps of sweep.cuh, but without a loader: the
// record stream comes page by page through cp.async (next page in flight under the current one), the already-solved
// rows of a block are read straight from global memory / L2 -- all of a block's rows requested before the first is
// used -- the tile lives in registers (16 rows, one column per lane) and ~32 warps per SM hide each other's latencies.
// The persistent kernel keeps the narrow levels at the top of the tree, where the critical path -- not throughput --
// decides.  Same records, same arithmetic per column as sweep.cuh (fast mode only: exact-order schedules stay with the
// persistent kernel).
// What it replaces in the reference: solve_lower_* / solve_upper_* for these rows (kernels/kernels.cu:64-131, 133-202).
#pragma once
#include <cstdint>

#include "kernel_args.h"
#include "local_sweep.cuh"
#include "sweep.cuh"

namespace wide_sweep {

using chb::SweepTask;
using chb::WideArgs;
using local_sweep::cp_async;
using local_sweep::cp_async_commit;
using local_sweep::cp_async_wait;
using sweep::fnma;
using sweep::kPage;
using sweep::Vec4;
using sweep::W;

template <typename F, bool UPPER>
__device__ __forceinline__ void wide_worker(const WideArgs<F> &a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // two pages of the task's record stream
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.k, kp = (k + 3) & ~3;
    const int col = warp * 32 + lane;
    const bool active = col < k;
    const int lcol = active ? col : 0;   // a column this lane may always read
    auto page_s = [&](int b) { return smem_raw + (size_t) b * kPage; };
    auto issue_page = [&](const SweepTask &td, uint32_t pg, int b) {
        const uint32_t bytes = min((uint32_t) kPage, td.bytes - pg * kPage);
        const unsigned char *src = a.slab + td.offset + (size_t) pg * kPage;
        for (uint32_t o = 16u * threadIdx.x; o < bytes; o += 16u * blockDim.x) cp_async<16>(page_s(b) + o, src + o);
    };

    int t = blockIdx.x;
    if (t >= a.n_tasks) return;
    SweepTask td = a.tasks[t];
    uint32_t pg = 0;
    int buf = 0;
    issue_page(td, 0, 0);
    cp_async_commit();

    F T[W];
    int row0 = 0, w = 0;
    for (;;) {
        const uint32_t n_pages = (td.bytes + kPage - 1) / kPage;
        // the page after this one (of this task or of the CTA's next task) goes in flight first
        int nt = t;
        uint32_t npg = pg + 1;
        SweepTask ntd = td;
        if (npg >= n_pages) {
            nt = t + gridDim.x;
            npg = 0;
            if (nt < a.n_tasks) ntd = a.tasks[nt];
        }
        if (nt < a.n_tasks) issue_page(ntd, npg, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();  // this page (copied by all threads) is in shared memory

        const unsigned char *p = page_s(buf);
        const unsigned char *pend = p + min((uint32_t) kPage, td.bytes - pg * kPage);
        while (p < pend) {
            const uint4 h = *reinterpret_cast<const uint4 *>(p);
            const uint32_t type = h.x & 0x3f;
            const unsigned char *pay = p + 16;
            p += (h.x >> 8) * 16;
            if (type == sweep::REC_BLOCK_D4) {
                // tile -= L_block * already-solved rows, read where they lie (work vector or residual vector): every
                // source row of the block is requested before the first one is used
                const uint32_t smask = h.z & 0xffff, tmask = h.z >> 16;
                const bool g0 = (tmask & 0x000f) != 0, g1 = (tmask & 0x00f0) != 0, g2 = (tmask & 0x0f00) != 0, g3 = (tmask & 0xf000) != 0;
                const int stride = 4 * ((int) g0 + (int) g1 + (int) g2 + (int) g3);
                const F *val = reinterpret_cast<const F *>(pay);
                const F *rows = ((h.x & sweep::REC_SRC_R_FLAG) ? a.rbuf : a.tmp) + (size_t) (int) h.y * kp + lcol;
                F X[W];
#pragma unroll
                for (int s = 0; s < W; ++s) X[s] = (smask >> s & 1) ? __ldcg(rows + (size_t) s * kp) : F(0);
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    const int s = UPPER ? W - 1 - q : q;  // sweep order: ascending rows forward, descending backward
                    if (smask >> s & 1) {
                        const F x = X[s];
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
            } else if (!UPPER && type == sweep::REC_ADD_ROWS) {
                // partial sums formed by the local tasks below: tile[t] += row, targets of tmask in ascending order
                const uint32_t tmask = h.z >> 16;
                const F *row = a.cbuf + (size_t) h.y * kp + lcol;
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    if (tmask >> i & 1) {
                        T[i] += __ldcg(row);
                        row += kp;
                    }
                }
            } else if (type == sweep::REC_BEGIN_RHS) {
                row0 = (int) h.y;
                w = (int) h.z;
                const int *perm = reinterpret_cast<const int *>(pay);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    if (i < w) T[i] = UPPER ? __ldcg(a.tmp + (size_t) (row0 + i) * kp + lcol) : __ldg(a.src + (size_t) perm[i] * a.src_ld + lcol);
                    else T[i] = F(0);
                }
            } else if (type == sweep::REC_END_DENSE || type == sweep::REC_END_STORE) {
                const int *perm = reinterpret_cast<const int *>(pay);
                const bool to_residual = type == sweep::REC_END_STORE && h.w == 1;
                if (type == sweep::REC_END_DENSE)
                    sweep::tri_dense<F, UPPER>(T, reinterpret_cast<const F *>(pay + (UPPER ? 4 * W : 0)), w == 1 ? 1 : (w + 3) & ~3);
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
            } else if (type == sweep::REC_BEGIN_ZERO) {
                row0 = (int) h.y;
                w = (int) h.z;
#pragma unroll
                for (int i = 0; i < W; ++i) T[i] = F(0);
            } else if (type == sweep::REC_BEGIN_SLOT || type == sweep::REC_ADD_SLOT) {
                if (type == sweep::REC_BEGIN_SLOT) {
                    row0 = (int) h.y;
                    w = (int) (h.z & 0xff);
                }
                const F *slot = a.scratch + (size_t) (h.z >> 8) * W * kp + lcol;
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    const F x = __ldcg(slot + (size_t) i * kp);
                    T[i] = type == sweep::REC_BEGIN_SLOT ? x : T[i] + x;
                }
            } else if (type == sweep::REC_END_SLOT) {
                if (active) {
                    F *slot = a.scratch + (size_t) (h.z >> 8) * W * kp + col;
#pragma unroll
                    for (int i = 0; i < W; ++i) __stcg(slot + (size_t) i * kp, T[i]);
                }
            }
            // REC_NOP: page padding
        }
        // later launches (the next levels, the persistent kernel) find the task published
        if (pg + 1 == n_pages && threadIdx.x == 0) a.done[td.level] = a.epoch;
        __syncthreads();  // every thread is done with this page before the next copies land in its buffer
        if (nt >= a.n_tasks) break;
        t = nt, pg = npg, td = ntd, buf ^= 1;
    }
    cp_async_wait<0>();
}

}  // namespace wide_sweep

#define CHB_WIDE(NAME, T, UP)                                                                   \
    extern "C" __global__ void __launch_bounds__(128) NAME(const chb::WideArgs<T> a) {          \
        wide_sweep::wide_worker<T, UP>(a);                                                      \
    }
CHB_WIDE(chb_wide_f32_lower, float, false)
CHB_WIDE(chb_wide_f32_upper, float, true)
CHB_WIDE(chb_wide_f64_lower, double, false)
CHB_WIDE(chb_wide_f64_upper, double, true)
