// Local-task kernels (sm_100a): the LOCAL tasks of a panel image (panel_schedule.h: REC_LOCAL_LOAD / REC_LOCAL_ROWS /
// REC_LOCAL_STORE -- complete subtrees at the bottom of the elimination tree, most rows of a mesh factor) solved by
// a plain grid instead of by the persistent sweep kernel (sweep.cuh).
//
// Why: in the forward sweep a local task depends on nothing, in the backward sweep nothing depends on it, so all of
// them can run as one embarrassingly parallel launch before (forward) / after (backward) the dependency-driven part.
// The per-task timeline of cfg4 (profiles/r02y_trace_cfg4_levels.txt) shows them taking 1.9 ms of the 6.8 ms forward
// and 2.6 ms of the 5.1 ms backward sweep inside the persistent kernel (8 us per task per CTA: descriptor, slab page,
// row staging and write-back round trips in sequence through one loader warp, and the index decode of every sparse
// row repeated by each of the four 32-column compute warps).
//
// Execution model: one WARP per (task, column slice), a lane owning CPL adjacent RHS columns (4 x f32 / 2 x f64 =
// one 16-byte vector at 128 columns, so a sparse row is decoded once for all of a task's columns); the task's record
// stream and its rows are brought into the warp's shared memory with cp.async (every row of a task in flight at once,
// the next task's records prefetched under the arithmetic); a lane only ever touches its own columns, so the
// substitution needs no intra-warp synchronisation; summation order, fused multiply-subtract and the pivot
// handling are those of sweep.cuh's REC_LOCAL_ROWS branch (results are bit-identical to the persistent kernel's).
// What it replaces in the reference: the rows of these subtrees in solve_lower_* / solve_upper_*
// (kernels/kernels.cu:64-131, 133-202).
#pragma once
#include <cstdint>

#include "kernel_args.h"
#include "sweep.cuh"

namespace local_sweep {

using chb::LocalArgs;
using chb::SweepTask;
using sweep::fnma;
using sweep::smem_u32;
using sweep::Vec4;

template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src) {
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// CPL adjacent columns of one row, moved as one vector
template <typename F, int CPL> struct Cols {
    F v[CPL];
    __device__ __forceinline__ static Cols load(const F *p) {
        Cols c;
        if (sizeof(F) * CPL == 16) {
            const uint4 q = *reinterpret_cast<const uint4 *>(p);
            *reinterpret_cast<uint4 *>(c.v) = q;
        } else if (sizeof(F) * CPL == 8) {
            const uint2 q = *reinterpret_cast<const uint2 *>(p);
            *reinterpret_cast<uint2 *>(c.v) = q;
        } else {
            c.v[0] = *p;
        }
        return c;
    }
    __device__ __forceinline__ void store(F *p) const {
        if (sizeof(F) * CPL == 16) *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(v);
        else if (sizeof(F) * CPL == 8) *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(v);
        else *p = v[0];
    }
    // streaming store to global memory (the work vector is read back by other SMs: keep it out of L1)
    __device__ __forceinline__ void store_cg(F *p) const {
        if (sizeof(F) * CPL == 16) __stcg(reinterpret_cast<uint4 *>(p), *reinterpret_cast<const uint4 *>(v));
        else if (sizeof(F) * CPL == 8) __stcg(reinterpret_cast<uint2 *>(p), *reinterpret_cast<const uint2 *>(v));
        else __stcg(p, v[0]);
    }
};

template <typename F, int CPL, bool UPPER, bool EXACT>
__device__ __forceinline__ void local_worker(const LocalArgs<F> &a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int RS = 32 * CPL;                 // elements per shared-memory row (the warp's column slice)
    constexpr int VB = (int) sizeof(F) * CPL;    // bytes a lane moves per row
    const int lane = threadIdx.x;
    const int k = a.k, kp = (k + 3) & ~3;
    const int n_slices = (k + RS - 1) / RS;
    auto slab_s = [&](int b) { return smem_raw + (size_t) b * (size_t) a.slab_cap; };  // two record buffers
    F *rows = reinterpret_cast<F *>(smem_raw + 2 * (size_t) a.slab_cap);
    F *mine = rows + lane * CPL;                 // this lane's columns of local row 0
    const long long items = (long long) a.n_tasks * n_slices;

    auto fetch_slab = [&](long long it, int buf) {
        if (it < items) {
            const SweepTask td = a.tasks[(int) (it / n_slices)];
            const uint32_t bytes = td.bytes & 0x0fffffffu;
            const unsigned char *src = a.slab + td.offset;
            for (uint32_t o = 16u * lane; o < bytes; o += 512u) cp_async<16>(slab_s(buf) + o, src + o);
        }
        cp_async_commit();
    };
    // one row of the slice: global -> shared, as one vector when the lane's columns all exist and the address is
    // aligned (vec), else element by element
    auto copy_in = [&](F *dst, const F *src, int n_valid, bool vec) {
        if (n_valid == CPL && vec) {
            cp_async<VB>(dst, src);
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                if (c < n_valid) cp_async<(int) sizeof(F)>(dst + c, src + c);
        }
    };
    auto copy_out = [&](F *dst, const Cols<F, CPL> &v, int n_valid, bool vec) {
        if (n_valid == CPL && vec) {
            v.store_cg(dst);
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                if (c < n_valid) __stcg(dst + c, v.v[c]);
        }
    };

    long long it = blockIdx.x;
    int buf = 0;
    fetch_slab(it, buf);
    for (; it < items; it += gridDim.x, buf ^= 1) {
        const int t = (int) (it / n_slices), slice = (int) (it - (long long) t * n_slices);
        const SweepTask td = a.tasks[t];
        const uint32_t bytes = td.bytes & 0x0fffffffu;
        const int col0 = slice * RS + lane * CPL;
        const int n_valid = min(max(k - col0, 0), CPL);
        cp_async_wait<0>();
        __syncwarp();  // the task's records (copied by all lanes) are in shared memory

        const unsigned char *p = slab_s(buf), *pend = p + bytes;
        int n_int;
        {
            // REC_LOCAL_LOAD: the task's first record
            const uint4 h = *reinterpret_cast<const uint4 *>(p);
            const int *list = reinterpret_cast<const int *>(p + 16);
            const int row_lo = (int) h.y;
            n_int = (int) (h.z & 0xffff);
            const int n_ext = (int) (h.z >> 16);
            if (!UPPER) {
                for (int i = 0; i < n_int; ++i)
                    copy_in(mine + (size_t) i * RS, a.src + (size_t) list[i] * a.src_ld + col0, n_valid, a.src_vec != 0);
            } else {
                for (int i = 0; i < n_int; ++i)
                    copy_in(mine + (size_t) i * RS, a.tmp + (size_t) (row_lo + i) * kp + col0, n_valid, true);
                for (int e = 0; e < n_ext; ++e)
                    copy_in(mine + (size_t) (n_int + e) * RS, a.tmp + (size_t) list[e] * kp + col0, n_valid, true);
            }
            cp_async_commit();
            p += (h.x >> 8) * 16;
        }
        fetch_slab(it + gridDim.x, buf ^ 1);  // the next task's records, under this task's arithmetic
        // forward: where the task's contribution rows go (REC_LOCAL_STORE, the task's last record, has the base)
        size_t contrib_base = 0;
        if (!UPPER) {
            for (const unsigned char *q = p; q < pend;) {
                const uint4 hq = *reinterpret_cast<const uint4 *>(q);
                if ((hq.x & 0x3f) == sweep::REC_LOCAL_STORE) contrib_base = hq.w;
                q += (hq.x >> 8) * 16;
            }
        }
        cp_async_wait<1>();                   // this task's rows (a lane reads only what it copied itself)

        while (p < pend) {
            const uint4 h = *reinterpret_cast<const uint4 *>(p);
            const uint32_t type = h.x & 0x3f;
            const unsigned char *pay = p + 16;
            p += (h.x >> 8) * 16;
            if (type == sweep::REC_LOCAL_ROWS) {
                const unsigned char *q = pay;
                for (int i = 0; i < (int) h.y; ++i) {
                    const uint32_t meta = *reinterpret_cast<const uint32_t *>(q);
                    const F pivot = *reinterpret_cast<const F *>(q + 8);
                    const int n4 = (int) (meta >> 8 & 0xff);
                    const unsigned char *idx = q + 16;
                    const F *val = reinterpret_cast<const F *>(q + 16 + ((4 * n4 + 15) & ~15));
                    F *dst = mine + (size_t) (meta & 0xff) * RS;
                    const bool contribution = (meta >> 16 & 1) != 0;
                    Cols<F, CPL> acc;
                    if (contribution) {  // accumulated from zero, never read again by this task: registers -> global
#pragma unroll
                        for (int c = 0; c < CPL; ++c) acc.v[c] = F(0);
                    } else {
                        acc = Cols<F, CPL>::load(dst);
                    }
                    for (int g = 0; g < n4; ++g) {
                        const uint32_t i4 = *reinterpret_cast<const uint32_t *>(idx + 4 * g);
                        const Vec4<F> vv = Vec4<F>::load(val + 4 * g);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const Cols<F, CPL> s = Cols<F, CPL>::load(mine + (size_t) (i4 >> (8 * j) & 0xff) * RS);
#pragma unroll
                            for (int c = 0; c < CPL; ++c) acc.v[c] = fnma(vv.x[j], s.v[c], acc.v[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc.v[c] = EXACT ? acc.v[c] / pivot : acc.v[c] * pivot;
                    if (contribution)
                        copy_out(a.cbuf + (contrib_base + (size_t) ((int) (meta & 0xff) - n_int)) * kp + col0, acc, n_valid, true);
                    else
                        acc.store(dst);
                    q = reinterpret_cast<const unsigned char *>(val + 4 * n4);
                }
            } else if (type == sweep::REC_LOCAL_STORE) {
                const int lo = (int) h.y;
                const int *perm = reinterpret_cast<const int *>(pay);
                for (int i = 0; i < n_int; ++i) {
                    const Cols<F, CPL> v = Cols<F, CPL>::load(mine + (size_t) i * RS);
                    copy_out(a.tmp + (size_t) (lo + i) * kp + col0, v, n_valid, true);
                    if (UPPER) copy_out(a.dst + (size_t) perm[i] * a.dst_ld + col0, v, n_valid, a.dst_vec != 0);
                }
            }
            // REC_NOP (page padding): nothing
        }
        // forward: consumers (the persistent kernel, launched after this one) find the task published
        if (!UPPER && slice == 0 && lane == 0) a.done[td.level] = a.epoch;
        __syncwarp();  // every lane is done with this buffer pair before the next copies land in it
    }
    cp_async_wait<0>();
}

}  // namespace local_sweep

#define CHB_LOCAL(NAME, T, CPL, UP, EX)                                                        \
    extern "C" __global__ void __launch_bounds__(32) NAME(const chb::LocalArgs<T> a) {         \
        local_sweep::local_worker<T, CPL, UP, EX>(a);                                          \
    }
#define CHB_LOCAL_SET(SUFFIX, T, CPL)                                    \
    CHB_LOCAL(chb_local_##SUFFIX##_lower, T, CPL, false, false)          \
    CHB_LOCAL(chb_local_##SUFFIX##_upper, T, CPL, true, false)           \
    CHB_LOCAL(chb_local_##SUFFIX##_lower_exact, T, CPL, false, true)     \
    CHB_LOCAL(chb_local_##SUFFIX##_upper_exact, T, CPL, true, true)

CHB_LOCAL_SET(f32_c1, float, 1)
CHB_LOCAL_SET(f32_c2, float, 2)
CHB_LOCAL_SET(f32_c4, float, 4)
CHB_LOCAL_SET(f64_c1, double, 1)
CHB_LOCAL_SET(f64_c2, double, 2)
