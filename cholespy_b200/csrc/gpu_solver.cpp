// Device-resident state and dispatch of the solve path.  See gpu_solver.h.
#include "gpu_solver.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <type_traits>

#include "disk_cache.h"
#include "kernel_args.h"

namespace chb {

namespace {

constexpr int kAnalysisBlock = 256;  // threads per CTA of the analysis kernel
constexpr int kSweepStages = 4;      // pages in the CTA's slab ring (sweep.cuh: kStages)
constexpr int kSweepRowBars = 16;    // sweep.cuh: kRowBars
constexpr int kSweepTaskQueue = 8;   // sweep.cuh: kTaskQueue
constexpr int kFuncAttrMaxDynamicSharedSize = 8;  // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES
constexpr int kFuncAttrSharedCarveout = 9;        // CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT (percent)

// Throws on failure: a null work buffer must never reach a kernel launch (it would fault the device and poison the
// primary context shared with the caller's framework).
void *dev_alloc(size_t bytes) {
    void *p = nullptr;
    const CUresult rc = cuMemAlloc(&p, std::max<size_t>(bytes, 16));
    if (rc != CUDA_SUCCESS || !p) {
        cuda_check(rc);
        throw std::runtime_error("out of device memory (cuMemAlloc of " + std::to_string(bytes) + " bytes failed)");
    }
    return p;
}
void dev_free(void *&p) {
    if (p) cuda_check(cuMemFree(p));
    p = nullptr;
}
void h2d(void *dst, const void *src, size_t bytes) {
    if (bytes) cuda_check(cuMemcpyAsync(dst, src, bytes, nullptr));
}
void d2h(void *dst, const void *src, size_t bytes) {
    if (bytes) cuda_check(cuMemcpyAsync(dst, src, bytes, nullptr));
}
void sync0() { cuda_check(cuStreamSynchronize(nullptr)); }

struct ContextGuard {
    explicit ContextGuard(DeviceState *d) { cuda_check(cuCtxPushCurrent(d->ctx)); }
    ~ContextGuard() { cuda_check(cuCtxPopCurrent(nullptr)); }
};

DeviceState *require_device(int device) {
    DeviceState *d = device_state(device);
    if (!d) throw std::runtime_error(std::string("CUDA backend unavailable: ") + driver_error_text());
    return d;
}

// Stable bucket of rows by level: ascending row id inside each level
// (behaviour of reference src/cholesky_solver.cpp:348-367).
void bucket_rows_by_level(const std::vector<int> &level_of_row, int n_levels, std::vector<int> &order) {
    const int n = (int) level_of_row.size();
    std::vector<int> cursor(n_levels + 1, 0);
    for (int r = 0; r < n; ++r) cursor[level_of_row[r] + 1]++;
    for (int l = 0; l < n_levels; ++l) cursor[l + 1] += cursor[l];
    order.resize(n);
    for (int r = 0; r < n; ++r) order[cursor[level_of_row[r]]++] = r;
}

// Pattern of the CSR form of a lower-triangular CSC operand (rows sorted by column, diagonal last).
void csc_to_csr_pattern(int n, const std::vector<int> &colptr, const std::vector<int> &rowidx, std::vector<int> &rowptr,
                        std::vector<int> &colidx) {
    const size_t nnz = (size_t) colptr[n];
    rowptr.assign(n + 1, 0);
    for (size_t p = 0; p < nnz; ++p) rowptr[rowidx[p] + 1]++;
    std::partial_sum(rowptr.begin(), rowptr.end(), rowptr.begin());
    colidx.resize(nnz);
    std::vector<int> cursor(rowptr.begin(), rowptr.end() - 1);
    for (int c = 0; c < n; ++c)
        for (int p = colptr[c]; p < colptr[c + 1]; ++p) colidx[cursor[rowidx[p]]++] = c;
}

int padded_rhs(int k) { return (k + 3) & ~3; }  // row stride of the work vector (whole 16-byte chunks)

int sweep_warps(int k) { return (k + 31) / 32; }  // compute warps: one per 32 RHS columns (+ 1 loader warp).  Wider lanes
// (2 or 4 columns per lane, fewer warps) were measured slower on ico8 and ico9 at 64..128 RHS: tasks are latency-bound.

// Rows of the shared-memory row ring: room for the largest local task the persistent kernel still executes (none by
// default: they run in the local-task kernel) plus one full block, at least 20 KB (f32: six CTAs per SM at 128
// right-hand sides) / 36 KB (f64: four); measured: occupancy beats ring depth.
int sweep_ring_rows(int elem, int k, int local_rows) {
    const int row_bytes = padded_rhs(k) * elem;
    // f32: six CTAs per SM at 128 columns (sweep.cuh: CHB_SWEEP_MIN_CTAS_F32) leave 20 KB next to the four slab pages
    int budget_kb = elem == 4 ? 20 : 36;
    if (const char *v = std::getenv("CHB_RING_KB")) budget_kb = std::max(1, std::atoi(v));
    int slack = 16;  // rows beyond the largest local task: one full block next to it (measured on cfg4: 16 beats 32, one more CTA per SM)
    if (const char *v = std::getenv("CHB_RING_SLACK")) slack = std::max(16, std::atoi(v));
    return std::max(local_rows + slack, std::min(256, (budget_kb << 10) / row_bytes));
}

// Dynamic shared memory of one sweep CTA (sweep.cuh: struct Smem).
size_t sweep_smem_bytes(int elem, int k, int local_rows) {
    const size_t row_bytes = (size_t) padded_rhs(k) * (size_t) elem;
    return kSweepStages * (size_t) page_bytes_for(elem) + (size_t) sweep_ring_rows(elem, k, local_rows) * row_bytes +
           (2 * kSweepStages + 2 * kSweepRowBars + 2 * kSweepTaskQueue) * 8 + kSweepTaskQueue * 32 + kSweepRowBars * 4;
}

// -- front kernels (front_sweep.cuh) --------------------------------------------------------------
constexpr int kFrontStages = 3;      // front_sweep.cuh: kStages
constexpr int kFrontRowBars = 16;    // kRowBars
constexpr int kFrontTaskQueue = 4;   // kTaskQueue
constexpr size_t kSmemPerSm = 233472, kSmemReservedPerCta = 1024;

int front_cpl(int k) {
    const int kp = padded_rhs(k);
    int cpl = kp <= 32 ? 1 : kp <= 64 ? 2 : 4;
    if (const char *v = std::getenv("CHB_FRONT_CPL")) {
        const int c = std::atoi(v);
        if ((c == 1 || c == 2 || c == 4) && c * 32 >= kp) cpl = c;
    }
    return cpl;
}
size_t front_fixed_smem(int page_bytes) {
    return (size_t) kFrontStages * (size_t) page_bytes + (2 * kFrontStages + 2 * kFrontRowBars + 2 * kFrontTaskQueue) * 8 +
           kFrontTaskQueue * 16 + 2 * kFrontRowBars * 4 + 64;
}
// Rows of the row ring: what the CTA shape's share of the SM's shared memory holds (2 CTAs per SM with 7 compute
// warps, 3 with 4), at most a few chunks beyond the largest staged record for narrow right-hand sides.
int front_ring_rows(int elem, int k, int page_bytes, int n_warps, int max_stage_rows) {
    const size_t row_bytes = (size_t) padded_rhs(k) * (size_t) elem;
    const int ctas = n_warps == 4 ? 3 : 2;
    const size_t budget = kSmemPerSm / ctas - kSmemReservedPerCta - front_fixed_smem(page_bytes);
    int rows = (int) (budget / row_bytes) - 1;
    if (const char *v = std::getenv("CHB_FRONT_RING_ROWS")) rows = std::min(rows, std::max(1, std::atoi(v)));
    rows = std::min(rows, std::max(256, max_stage_rows + 4 * kFrontLoad));
    if (rows < max_stage_rows) throw std::runtime_error("front sweep: the largest staged record does not fit the shared-memory row ring");
    return rows;
}
size_t front_smem_bytes(int elem, int k, int page_bytes, int ring_rows) {
    return front_fixed_smem(page_bytes) + (size_t) (ring_rows + 1) * (size_t) padded_rhs(k) * (size_t) elem;
}

}  // namespace

// Flag bits of SweepTask::bytes in the replicated (ticket-ordered) task table of a panel image; the tables the
// kernels walk (split_tasks) carry plain byte counts (+ the ticket-batch hint in bits 28-30 for the persistent kernel).
constexpr uint32_t kTaskBytesMask = 0x03ffffffu;    // bits 0-25: bytes of the task's record stream
constexpr uint32_t kTaskLocalKept = 1u << 26;       // a local task left to the persistent kernel (too large for the local-task kernel)
constexpr uint32_t kTaskLevelParity = 1u << 27;     // parity of the task's dependency level (level boundaries in ticket order)
constexpr uint32_t kTaskLocalKernel = 1u << 31;     // a local task the local-task kernel executes

void ImageShape::to_words(int64_t *w) const {
    w[0] = n, w[1] = nnz;
    for (int u = 0; u < 2; ++u) {
        w[2 + u] = n_tasks[u], w[4 + u] = (int64_t) slab_bytes[u], w[6 + u] = n_slots[u], w[8 + u] = n_task_levels[u];
        w[10 + u] = n_levels[u];
        w[12 + u] = max_local_rows[u];
    }
    w[14] = exact_order;
    w[15] = n_contrib_rows;
    w[16] = front;
    for (int u = 0; u < 2; ++u) w[17 + u] = n_w_rows[u], w[19 + u] = max_stage_rows[u];
    w[21] = n_warps, w[22] = page_bytes, w[23] = local_rows_cap[0] | local_rows_cap[1] << 16 | main_local_rows[0] << 32 | main_local_rows[1] << 48;
}
ImageShape ImageShape::from_words(const int64_t *w) {
    ImageShape s;
    s.n = w[0], s.nnz = w[1];
    for (int u = 0; u < 2; ++u) {
        s.n_tasks[u] = w[2 + u], s.slab_bytes[u] = w[4 + u], s.n_slots[u] = w[6 + u], s.n_task_levels[u] = w[8 + u];
        s.n_levels[u] = w[10 + u];
        s.max_local_rows[u] = w[12 + u];
    }
    s.exact_order = w[14];
    s.n_contrib_rows = w[15];
    s.front = w[16];
    for (int u = 0; u < 2; ++u) s.n_w_rows[u] = w[17 + u], s.max_stage_rows[u] = w[19 + u];
    s.n_warps = w[21], s.page_bytes = w[22];
    s.local_rows_cap[0] = w[23] & 0xffff, s.local_rows_cap[1] = w[23] >> 16 & 0xffff;
    s.main_local_rows[0] = w[23] >> 32 & 0xffff, s.main_local_rows[1] = w[23] >> 48 & 0x7fff;
    return s;
}

void GpuSolver::allocate_sweep(SweepDev &s) {
    s.slab = dev_alloc(s.slab_bytes);
    s.tasks = dev_alloc(sizeof(SweepTask) * (size_t) s.n_tasks);
    s.done = dev_alloc(sizeof(int) * (size_t) s.n_tasks);
    cuda_check(cuMemsetD32Async(s.done, 0, (size_t) std::max(s.n_tasks, 1), nullptr));
    const char *trace = std::getenv("CHB_TRACE");
    if (trace && std::atoi(trace) != 0) {
        // 8 words per task + 16 words per CTA (time spent by compute warp 0 per record type, front kernels)
        const size_t words = 8 * (size_t) s.n_tasks + 16 * 4096;
        s.trace = dev_alloc(sizeof(unsigned long long) * words);
        cuda_check(cuMemsetD8Async(s.trace, 0, sizeof(unsigned long long) * words, nullptr));
    }
}

size_t GpuSolver::copy_trace(bool upper, unsigned long long *out, size_t cap) {
    const SweepDev &s = m_sweep[upper ? 1 : 0];
    if (!s.trace) return 0;
    ContextGuard guard(m_dev);
    const size_t words = std::min(cap, (size_t) 8 * (size_t) s.n_tasks + 16 * 4096);
    d2h(out, s.trace, sizeof(unsigned long long) * words);
    sync0();
    return words;
}

void GpuSolver::upload_sweep(SweepDev &d, const SweepImage &img) {
    static_assert(sizeof(SweepTask) == sizeof(TaskDesc), "host and device task descriptors must match");
    d.n_tasks = (int) img.tasks.size();
    d.slab_bytes = img.slab.size();
    d.n_slots = img.n_slots;
    d.n_task_levels = img.n_levels;
    d.max_local_rows = img.max_local_rows;
    d.n_contrib_rows = img.n_contrib_rows;
    d.n_stored_values = img.n_stored_values;
    allocate_sweep(d);
    // device task table in ticket order; the `level` word carries the task id (its flag index)
    std::vector<TaskDesc> by_ticket(img.tasks.size());
    // ... and bits 28-30 of `bytes` say how many tickets a loader may draw together at this position (sweep.cuh):
    // 8 while at least kWideLevel tasks of the same dependency level are still to be handed out, else 1
    size_t kWideLevel = 65536;  // (re-measured with 6 CTAs per SM, profiles/r02x_inv_ab.log: batching tickets rarely pays)
    if (const char *v = std::getenv("CHB_WIDE_LEVEL")) kWideLevel = (size_t) std::max(1, std::atoi(v));
    // selection rule of the local-task kernel: every local task by default; CHB_LOCAL_KERNEL=0 leaves them all to the
    // persistent kernel (the round-1 behaviour); CHB_LOCAL_KERNEL_ROWS="fwd,bwd" / CHB_LOCAL_KERNEL_BYTES leave
    // the ones holding more rows / records than that to it (tuning)
    const bool upper = &d == &m_sweep[1];
    bool use_local = true;
    int rows_cap = 255, bytes_cap = 32768, rows_max = 0;
    if (const char *v = std::getenv("CHB_LOCAL_KERNEL")) use_local = std::atoi(v) != 0;
    if (const char *v = std::getenv("CHB_LOCAL_KERNEL_ROWS")) {
        int f = 0, b = 0;
        if (std::sscanf(v, "%d%*[,:]%d", &f, &b) == 2) rows_cap = std::max(1, std::min(255, upper ? b : f));
    }
    if (const char *v = std::getenv("CHB_LOCAL_KERNEL_BYTES")) bytes_cap = std::max(256, std::min(32768, std::atoi(v)));
    const size_t nt = img.order.size();
    size_t level_end = 0;
    for (size_t t = 0; t < nt; ++t) {
        const TaskDesc &src = img.tasks[(size_t) img.order[t]];
        if (t >= level_end) {
            level_end = t + 1;
            while (level_end < nt && img.tasks[(size_t) img.order[level_end]].level == src.level) ++level_end;
        }
        if (src.bytes > kTaskBytesMask) throw std::logic_error("sweep image: task slab too large");
        by_ticket[t] = src;
        by_ticket[t].level = (uint32_t) img.order[t];
        by_ticket[t].bytes |= (level_end - t >= kWideLevel ? 7u : 0u) << 28;
        // bit 31: a local task the local-task kernel takes (local_sweep.cuh): one whose rows and records fit that
        // kernel's shared-memory budget per warp; the few larger ones stay with the persistent kernel
        uint32_t head[4] = {0, 0, 0, 0};
        std::memcpy(head, img.slab.data() + src.offset, sizeof head);
        if (use_local && (head[0] & 0x3fu) == (uint32_t) REC_LOCAL_LOAD) {
            const int n_int = (int) (head[2] & 0xffffu), n_ext = (int) (head[2] >> 16);
            const int rows = upper ? n_int + n_ext : n_int;   // forward contributions go straight to global memory
            if (rows <= rows_cap && (int) src.bytes <= bytes_cap) by_ticket[t].bytes |= kTaskLocalKernel, rows_max = std::max(rows_max, rows);
            else by_ticket[t].bytes |= kTaskLocalKept, d.main_local_rows = std::max(d.main_local_rows, n_int + n_ext);
        } else if ((head[0] & 0x3fu) == (uint32_t) REC_LOCAL_LOAD) {
            by_ticket[t].bytes |= kTaskLocalKept;
            d.main_local_rows = std::max(d.main_local_rows, (int) (head[2] & 0xffffu) + (int) (head[2] >> 16));
        }
        if (src.level & 1u) by_ticket[t].bytes |= kTaskLevelParity;
    }
    d.local_rows_cap = rows_max;
    h2d(d.slab, img.slab.data(), img.slab.size());
    h2d(d.tasks, by_ticket.data(), sizeof(TaskDesc) * by_ticket.size());
    sync0();
    split_tasks(d, by_ticket.data());
}

// The task tables of the three executors of a panel sweep, derived from the replicated one (replicas receive
// `tasks` from a peer and derive theirs on first use):
//   local  -- local tasks taken by the local-task kernel (bit 31),
//   wide   -- the panel tasks of the wide dependency levels next to the leaves, level by level (plain launches),
//   main   -- everything else: the persistent kernel.
// Wide levels (opt-in, CHB_WIDE_MIN_TASKS=N > 0): forward, every level up to the last one holding >= N such tasks;
// backward, every level from the first such one on.  Exact-order images, or N = 0 (the default): none.
void GpuSolver::split_tasks(SweepDev &s, const void *host_tasks) {
    const bool upper = &s == &m_sweep[1];
    std::vector<TaskDesc> all((size_t) s.n_tasks);
    if (host_tasks) {
        std::memcpy(all.data(), host_tasks, sizeof(TaskDesc) * all.size());
    } else {
        d2h(all.data(), s.tasks, sizeof(TaskDesc) * all.size());
        sync0();
    }
    // dependency level of every task (ticket order is level order; bit 27 = the level's parity) and the number of
    // wide-kernel candidates per level
    std::vector<int> level_of(all.size(), 0);
    std::vector<int> candidates;
    int level = 0;
    for (size_t t = 0; t < all.size(); ++t) {
        if (t > 0 && ((all[t].bytes ^ all[t - 1].bytes) & kTaskLevelParity)) level += 1;
        level_of[t] = level;
        if ((size_t) level >= candidates.size()) candidates.resize((size_t) level + 1, 0);
        if (!(all[t].bytes & (kTaskLocalKernel | kTaskLocalKept))) candidates[(size_t) level] += 1;
    }
    int min_tasks = 0;  // measured (profiles/r02x_wide_ab2.log): no gain over the persistent kernel on cfg4 / cfg5 at any threshold; opt-in
    if (const char *v = std::getenv("CHB_WIDE_MIN_TASKS")) min_tasks = std::atoi(v);
    int wide_lo = 0, wide_hi = -1;  // levels [wide_lo, wide_hi] go to the wide kernel
    if (min_tasks > 0 && !m_exact && !all.empty()) {
        const int n_levels = (int) candidates.size();
        if (!upper) {
            for (int l = 0; l < n_levels; ++l)
                if (candidates[(size_t) l] >= min_tasks) wide_hi = l;
            // a wide level must not read what a local task left to the persistent kernel (which runs last) produces
            for (size_t t = 0; t < all.size(); ++t)
                if (all[t].bytes & kTaskLocalKept) wide_hi = std::min(wide_hi, level_of[t] - 1);
        } else {
            for (int l = n_levels - 1; l >= 0; --l)
                if (candidates[(size_t) l] >= min_tasks) wide_lo = l, wide_hi = n_levels - 1;
            // a local task left to the persistent kernel (which runs first) must not wait for a wide level
            for (size_t t = 0; t < all.size(); ++t)
                if (all[t].bytes & kTaskLocalKept) wide_lo = std::max(wide_lo, level_of[t] + 1);
        }
    }
    std::vector<TaskDesc> main_t, local_t, wide_t;
    s.wide_levels.clear();
    uint32_t cap = 0;
    int cur_level = -1;
    for (size_t t = 0; t < all.size(); ++t) {
        TaskDesc d = all[t];
        const uint32_t flags = d.bytes;
        d.bytes &= ~(kTaskLocalKernel | kTaskLocalKept | kTaskLevelParity);
        if (flags & kTaskLocalKernel) {
            d.bytes &= kTaskBytesMask;
            local_t.push_back(d);
            cap = std::max(cap, d.bytes);
        } else if (!(flags & kTaskLocalKept) && level_of[t] >= wide_lo && level_of[t] <= wide_hi) {
            d.bytes &= kTaskBytesMask;
            if (level_of[t] != cur_level) {
                cur_level = level_of[t];
                s.wide_levels.push_back({(int) wide_t.size(), 0});
            }
            s.wide_levels.back().second += 1;
            wide_t.push_back(d);
        } else {
            main_t.push_back(d);  // keeps the ticket-batch hint (bits 28-30)
        }
    }
    s.n_main = (int) main_t.size();
    s.n_local = (int) local_t.size();
    s.n_wide = (int) wide_t.size();
    s.local_slab_cap = (int) ((cap + 15u) & ~15u);
    s.main_tasks = dev_alloc(sizeof(TaskDesc) * std::max<size_t>(main_t.size(), 1));
    s.local_tasks = dev_alloc(sizeof(TaskDesc) * std::max<size_t>(local_t.size(), 1));
    s.wide_tasks = dev_alloc(sizeof(TaskDesc) * std::max<size_t>(wide_t.size(), 1));
    if (!main_t.empty()) h2d(s.main_tasks, main_t.data(), sizeof(TaskDesc) * main_t.size());
    if (!local_t.empty()) h2d(s.local_tasks, local_t.data(), sizeof(TaskDesc) * local_t.size());
    if (!wide_t.empty()) h2d(s.wide_tasks, wide_t.data(), sizeof(TaskDesc) * wide_t.size());
    sync0();
    s.split = true;
}

void GpuSolver::upload_front(SweepDev &d, const FrontImage &img) {
    static_assert(sizeof(SweepTask) == sizeof(TaskDesc), "host and device task descriptors must match");
    d.n_tasks = (int) img.tasks.size();
    d.slab_bytes = img.slab.size();
    d.n_task_levels = img.n_levels;
    d.n_stored_values = img.n_stored_values;
    d.n_w_rows = img.n_w_rows;
    d.max_stage_rows = img.max_stage_rows;
    d.page_bytes = img.page_bytes;
    d.n_warps = img.n_warps;
    allocate_sweep(d);
    // device task table in ticket order; the `level` word carries the task id (its flag index)
    std::vector<TaskDesc> by_ticket(img.tasks.size());
    for (size_t t = 0; t < img.order.size(); ++t) {
        by_ticket[t] = img.tasks[(size_t) img.order[t]];
        by_ticket[t].level = (uint32_t) img.order[t];
    }
    h2d(d.slab, img.slab.data(), img.slab.size());
    h2d(d.tasks, by_ticket.data(), sizeof(TaskDesc) * by_ticket.size());
    sync0();
}

void GpuSolver::analyze(TriangleAnalysis &t, const std::vector<int> &ptr, const std::vector<int> &idx, bool lower) {
    int *counters = (int *) m_counters;
    void *rowptr_d = dev_alloc(sizeof(int) * ptr.size());
    void *colidx_d = dev_alloc(sizeof(int) * idx.size());
    void *level_d = dev_alloc(sizeof(int) * (size_t) m_n);
    h2d(rowptr_d, ptr.data(), sizeof(int) * ptr.size());
    h2d(colidx_d, idx.data(), sizeof(int) * idx.size());
    cuda_check(cuMemsetD8Async(level_d, 0xff, sizeof(int) * (size_t) m_n, nullptr));  // all -1
    cuda_check(cuMemsetD32Async(counters + 2, 0, 2, nullptr));

    AnalysisArgs args;
    args.n = m_n;
    args.descending = lower ? 0 : 1;
    args.rowptr = (const int *) rowptr_d;
    args.colidx = (const int *) colidx_d;
    args.level = (int *) level_d;
    args.max_level = counters + 3;
    args.ticket = counters + 2;
    CUfunction fn = kernel(m_dev, "chb_analysis");
    if (!fn) throw std::runtime_error(driver_error_text());
    const int grid = std::max(1, max_resident_blocks(m_dev, fn, kAnalysisBlock, 0));
    void *params[] = {&args};
    cuda_check(cuLaunchCooperativeKernel(fn, grid, 1, 1, kAnalysisBlock, 1, 1, 0, nullptr, params));

    t.level_of_row.resize(m_n);
    int max_level = 0;
    d2h(t.level_of_row.data(), level_d, sizeof(int) * (size_t) m_n);
    d2h(&max_level, counters + 3, sizeof(int));
    sync0();
    dev_free(level_d);
    dev_free(rowptr_d);
    dev_free(colidx_d);
    t.n_levels = max_level + 1;
    for (int r = 0; r < m_n; ++r)
        if (t.level_of_row[r] < 0 || t.level_of_row[r] > max_level)
            throw std::runtime_error("level analysis produced an out-of-range level (GPU kernel failure)");
    bucket_rows_by_level(t.level_of_row, t.n_levels, t.order_host);
}

// Both level arrays from the elimination tree (kernels.cu: chb_etree_round): forward level = height of the node,
// backward level = depth.  The parent array (first below-diagonal row of every column of L) is read off the factor
// on the host in O(n); the device doubles ancestor pointers, ceil(log2(height)) + 1 rounds of n threads.
void GpuSolver::analyze_etree(const HostFactor &L) {
    const int n = m_n;
    std::vector<int> parent((size_t) n), ones((size_t) n);
    for (int j = 0; j < n; ++j) {
        parent[(size_t) j] = L.colptr[j + 1] - L.colptr[j] > 1 ? L.rowidx[(size_t) L.colptr[j] + 1] : -1;
        ones[(size_t) j] = parent[(size_t) j] >= 0 ? 1 : 0;
    }
    const size_t bytes = sizeof(int) * (size_t) n;
    void *buf[6];  // anc, dist, best: two copies each (a round reads one set and writes the other)
    for (void *&b : buf) b = dev_alloc(bytes);
    int *counters = (int *) m_counters;
    h2d(buf[0], parent.data(), bytes);
    h2d(buf[2], ones.data(), bytes);
    cuda_check(cuMemsetD8Async(buf[4], 0, bytes, nullptr));
    CUfunction fn = kernel(m_dev, "chb_etree_round");
    if (!fn) throw std::runtime_error(driver_error_text());
    int cur = 0, active = n > 0 ? 1 : 0;
    for (int round = 0; round < 32 && active > 0; ++round) {
        EtreeArgs args;
        args.n = n;
        args.step = 1 << round;
        args.anc_in = (const int *) buf[0 + cur], args.anc_out = (int *) buf[0 + (cur ^ 1)];
        args.dist_in = (const int *) buf[2 + cur], args.dist_out = (int *) buf[2 + (cur ^ 1)];
        args.best_in = (const int *) buf[4 + cur], args.best_out = (int *) buf[4 + (cur ^ 1)];
        args.active = counters + 2;
        cuda_check(cuMemsetD8Async(args.best_out, 0, bytes, nullptr));
        cuda_check(cuMemsetD32Async(counters + 2, 0, 1, nullptr));
        void *params[] = {&args};
        cuda_check(cuLaunchKernel(fn, (unsigned) ((n + 255) / 256), 1, 1, 256, 1, 1, 0, nullptr, params, nullptr));
        d2h(&active, counters + 2, sizeof(int));
        sync0();
        cur ^= 1;
    }
    if (active > 0) throw std::runtime_error("elimination-tree analysis did not converge (the factor's tree is deeper than 2^32?)");
    m_lower.level_of_row.resize((size_t) n);
    m_upper.level_of_row.resize((size_t) n);
    d2h(m_lower.level_of_row.data(), buf[4 + cur], bytes);  // height
    d2h(m_upper.level_of_row.data(), buf[2 + cur], bytes);  // depth
    sync0();
    for (void *&b : buf) dev_free(b);
    for (TriangleAnalysis *t : {&m_lower, &m_upper}) {
        int max_level = 0;
        for (int r = 0; r < n; ++r) {
            if (t->level_of_row[(size_t) r] < 0 || t->level_of_row[(size_t) r] >= n)
                throw std::runtime_error("level analysis produced an out-of-range level (GPU kernel failure)");
            max_level = std::max(max_level, t->level_of_row[(size_t) r]);
        }
        t->n_levels = n > 0 ? max_level + 1 : 0;
        bucket_rows_by_level(t->level_of_row, t->n_levels, t->order_host);
    }
}

GpuSolver::GpuSolver(int device, int elem_size, const HostFactor &L)
    : m_dev(require_device(device)), m_elem(elem_size), m_n(L.n), m_nnz(L.nnz()) {
    ContextGuard guard(m_dev);
    m_counters = dev_alloc(sizeof(int) * 8);
    cuda_check(cuMemsetD32Async(m_counters, 0, 8, nullptr));
    m_perm_host = L.perm;

    // Host: lay the factor out as the two sweep images (one thread each) while the device runs the
    // level analysis of both triangles.  Default: the panel image (panel_schedule.h), the faster of the two on every
    // measured workload; CHB_SWEEP=front selects the supernodal front image (front_schedule.h).
    // CHB_EXACT_ORDER=1 (reference summation order) exists for the panel image only.
    const ScheduleOptions opt = schedule_options_from_env();
    m_exact = opt.exact_order;
    const char *mode = std::getenv("CHB_SWEEP");
    m_front = !m_exact && mode && std::strcmp(mode, "front") == 0;
    const FrontOptions fopt = front_options_from_env(m_elem);
    SweepImage img[2];
    FrontImage fimg[2];
    // CHB_CACHE_DIR set: images built before for this factor, element size and options are read back
    const bool caching = !cache_dir().empty();
    CacheKey key;
    std::string tag;
    bool cached = false;
    if (caching) {
        key = factor_key(L);
        char buf[256];
        if (m_front)
            snprintf(buf, sizeof buf, "front e%d w%d lr%d le%d p%d f%d s%d o%d b%d", m_elem, fopt.warps, fopt.local_rows, fopt.local_ext,
                     fopt.panel, fopt.fuse_steps, fopt.split_chunks, fopt.own_chunks, fopt.order_blocks);
        else
            snprintf(buf, sizeof buf, "panel e%d x%d lr%d le%d lc%d sc%d cc%d a%.4f b%.4f c%d i%d j%d", m_elem, (int) opt.exact_order, opt.local_rows,
                     opt.local_ext, opt.local_cost, opt.split_cost, opt.chunk_cost, opt.relax_alpha, opt.relax_beta, (int) opt.contributions,
                     opt.inv_panels, opt.inv_chunk);
        tag = buf;
        cached = m_front ? load_front_images(key, tag, &fimg[0], &fimg[1]) : load_panel_images(key, tag, &img[0], &img[1]);
    }
    PanelStructure structure;
    FrontStructure fstructure;
    if (!cached) {
        if (m_front) build_front_structure(L, m_elem, fopt, &fstructure);
        else build_panel_structure(L, opt, &structure);
    }
    std::exception_ptr errors[2];
    std::thread builders[2];
    for (int u = 0; u < 2; ++u)
        builders[u] = std::thread([&, u] {
            try {
                if (cached) return;
                if (m_front) build_front_image(L, fstructure, m_elem, u == 1, fopt, &fimg[u]);
                else build_sweep_image(L, structure, m_elem, u == 1, opt, &img[u]);
            } catch (...) {
                errors[u] = std::current_exception();
            }
        });
    try {
        // CHB_ANALYSIS=rows: the row-by-row restatement of the reference's kernels (one warp per row, ready flags),
        // kept as a cross-check; default: both level arrays from the elimination tree by pointer doubling
        const char *amode = std::getenv("CHB_ANALYSIS");
        if (amode && std::strcmp(amode, "rows") == 0) {
            {
                std::vector<int> rowptr, colidx;  // CSR of L (diagonal last per row): the forward operand
                csc_to_csr_pattern(L.n, L.colptr, L.rowidx, rowptr, colidx);
                analyze(m_lower, rowptr, colidx, true);
            }
            analyze(m_upper, L.colptr, L.rowidx, false);  // CSC of L == CSR of L' (diagonal first per row)
        } else {
            analyze_etree(L);
        }
    } catch (...) {
        for (auto &t : builders) t.join();
        throw;
    }
    for (auto &t : builders) t.join();
    for (int u = 0; u < 2; ++u)
        if (errors[u]) std::rethrow_exception(errors[u]);
    if (caching && !cached) {
        if (m_front) store_front_images(key, tag, fimg[0], fimg[1]);
        else store_panel_images(key, tag, img[0], img[1]);
    }
    for (int u = 0; u < 2; ++u) {
        if (m_front) {
            upload_front(m_sweep[u], fimg[u]);
            fimg[u] = FrontImage();
        } else {
            upload_sweep(m_sweep[u], img[u]);
            img[u] = SweepImage();
        }
    }
    if (cuda_first_error() != CUDA_SUCCESS) throw std::runtime_error(driver_error_text());
}

GpuSolver::GpuSolver(int device, int elem_size, const ImageShape &shape)
    : m_dev(require_device(device)), m_elem(elem_size), m_n((int) shape.n), m_nnz(shape.nnz) {
    ContextGuard guard(m_dev);
    m_counters = dev_alloc(sizeof(int) * 8);
    cuda_check(cuMemsetD32Async(m_counters, 0, 8, nullptr));
    for (int u = 0; u < 2; ++u) {
        m_sweep[u].n_tasks = (int) shape.n_tasks[u];
        m_sweep[u].slab_bytes = (uint64_t) shape.slab_bytes[u];
        m_sweep[u].n_slots = (int) shape.n_slots[u];
        m_sweep[u].n_task_levels = (int) shape.n_task_levels[u];
        m_sweep[u].max_local_rows = (int) shape.max_local_rows[u];
        m_sweep[u].local_rows_cap = (int) shape.local_rows_cap[u];
        m_sweep[u].main_local_rows = (int) shape.main_local_rows[u];
        m_sweep[u].n_w_rows = shape.n_w_rows[u];
        m_sweep[u].max_stage_rows = (int) shape.max_stage_rows[u];
        m_sweep[u].page_bytes = (int) shape.page_bytes;
        m_sweep[u].n_warps = (int) shape.n_warps;
        allocate_sweep(m_sweep[u]);
    }
    m_exact = shape.exact_order != 0;
    m_front = shape.front != 0;
    m_sweep[0].n_contrib_rows = shape.n_contrib_rows;
    m_lower.n_levels = (int) shape.n_levels[0];
    m_upper.n_levels = (int) shape.n_levels[1];
    sync0();
    if (cuda_first_error() != CUDA_SUCCESS) throw std::runtime_error(driver_error_text());
}

GpuSolver::~GpuSolver() {
    ContextGuard guard(m_dev);
    sync0();
    for (SweepDev &s : m_sweep) {
        dev_free(s.slab);
        dev_free(s.tasks);
        dev_free(s.done);
        dev_free(s.trace);
        dev_free(s.main_tasks);
        dev_free(s.local_tasks);
        dev_free(s.wide_tasks);
    }
    dev_free(m_counters);
    dev_free(m_tmp);
    dev_free(m_scratch);
    dev_free(m_rbuf);
    dev_free(m_cbuf);
    dev_free(m_wbuf);
    dev_free(m_stage_b);
    dev_free(m_stage_x);
    if (m_copy_in) {
        cuda_check(cuStreamDestroy(m_copy_in));
        cuda_check(cuStreamDestroy(m_copy_out));
        for (size_t i = 0; i < m_ev_in.size(); ++i) {
            cuda_check(cuEventDestroy(m_ev_in[i]));
            cuda_check(cuEventDestroy(m_ev_out[i]));
        }
    }
}

ImageShape GpuSolver::shape() const {
    ImageShape s;
    s.n = m_n, s.nnz = m_nnz;
    for (int u = 0; u < 2; ++u) {
        s.n_tasks[u] = m_sweep[u].n_tasks;
        s.slab_bytes[u] = (int64_t) m_sweep[u].slab_bytes;
        s.n_slots[u] = m_sweep[u].n_slots;
        s.n_task_levels[u] = m_sweep[u].n_task_levels;
        s.max_local_rows[u] = m_sweep[u].max_local_rows;
        s.local_rows_cap[u] = m_sweep[u].local_rows_cap;
        s.main_local_rows[u] = m_sweep[u].main_local_rows;
    }
    s.exact_order = m_exact ? 1 : 0;
    s.n_contrib_rows = m_sweep[0].n_contrib_rows;
    s.front = m_front ? 1 : 0;
    for (int u = 0; u < 2; ++u) s.n_w_rows[u] = m_sweep[u].n_w_rows, s.max_stage_rows[u] = m_sweep[u].max_stage_rows;
    s.n_warps = m_sweep[0].n_warps, s.page_bytes = m_sweep[0].page_bytes;
    s.n_levels[0] = m_lower.n_levels;
    s.n_levels[1] = m_upper.n_levels;
    return s;
}

std::vector<ImageBuffer> GpuSolver::image() const {
    std::vector<ImageBuffer> out;
    static const char *names[2][2] = {{"lower.slab", "lower.tasks"}, {"upper.slab", "upper.tasks"}};
    for (int u = 0; u < 2; ++u) {
        const SweepDev &s = m_sweep[u];
        out.push_back({names[u][0], s.slab, (size_t) s.slab_bytes});
        out.push_back({names[u][1], s.tasks, sizeof(SweepTask) * (size_t) s.n_tasks});
    }
    return out;
}

// columns a lane of the local-task kernel owns: the smallest of 1, 2, 4 (f64: 1, 2) that covers k with one warp
static int local_cpl(int elem, int k) {
    const int kp = (k + 3) & ~3, cap = elem == 4 ? 4 : 2;
    int cpl = 1;
    while (cpl < cap && 32 * cpl < kp) cpl *= 2;
    return cpl;
}

// The local tasks of one sweep as a plain launch (local_sweep.cuh).  The selection rule (upload_sweep) bounds the
// shared memory a warp needs: 255 rows x 512 B + 2 x 32 KB of records at most.
void GpuSolver::local_sweep(bool upper, int k, const void *src, void *dst, int ld) {
    const SweepDev &s = m_sweep[upper ? 1 : 0];
    if (s.n_local == 0) return;
    const int cpl = local_cpl(m_elem, k);
    const size_t smem = 2 * (size_t) s.local_slab_cap + (size_t) s.local_rows_cap * 32 * (size_t) cpl * (size_t) m_elem;
    char name[64];
    snprintf(name, sizeof name, "chb_local_%s_c%d_%s%s", m_elem == 4 ? "f32" : "f64", cpl, upper ? "upper" : "lower", m_exact ? "_exact" : "");
    CUfunction fn = kernel(m_dev, name);
    if (!fn) throw std::runtime_error(driver_error_text());
    cuda_check(cuFuncSetAttribute(fn, kFuncAttrMaxDynamicSharedSize, (int) smem));
    cuda_check(cuFuncSetAttribute(fn, kFuncAttrSharedCarveout, 100));
    const size_t vec_bytes = (size_t) cpl * (size_t) m_elem;
    auto fill = [&](auto &a) {
        using A = std::remove_reference_t<decltype(a)>;
        using Float = std::remove_pointer_t<decltype(A::tmp)>;
        a.n_tasks = s.n_local;
        a.k = k;
        a.epoch = m_epoch;
        a.slab_cap = s.local_slab_cap;
        a.src_vec = (((uintptr_t) src % vec_bytes) == 0 && ((size_t) ld * (size_t) m_elem) % vec_bytes == 0) ? 1 : 0;
        a.dst_vec = (((uintptr_t) dst % vec_bytes) == 0 && ((size_t) ld * (size_t) m_elem) % vec_bytes == 0) ? 1 : 0;
        a.src_ld = ld;
        a.dst_ld = ld;
        a.slab = (const unsigned char *) s.slab;
        a.tasks = (const SweepTask *) s.local_tasks;
        a.done = (int *) s.done;
        a.tmp = (Float *) m_tmp;
        a.src = (const Float *) src;
        a.dst = (Float *) dst;
        a.cbuf = (Float *) m_cbuf;
    };
    LocalArgs<float> a32;
    LocalArgs<double> a64;
    void *params[1];
    if (m_elem == 4) {
        fill(a32);
        params[0] = &a32;
    } else {
        fill(a64);
        params[0] = &a64;
    }
    const int n_slices = (k + 32 * cpl - 1) / (32 * cpl);
    const long long items = (long long) s.n_local * n_slices;
    int per_sm = max_resident_blocks(m_dev, fn, 32, smem) / std::max(1, m_dev->sm_count);
    if (const char *cap = std::getenv("CHB_LOCAL_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, std::atoi(cap)));
    const int grid = (int) std::max<long long>(1, std::min<long long>(items, (long long) per_sm * m_dev->sm_count));
    cuda_check(cuLaunchKernel(fn, (unsigned) grid, 1, 1, 32, 1, 1, (unsigned) smem, nullptr, params, nullptr));
}

// The wide dependency levels of one sweep, one plain launch per level (wide_sweep.cuh), in sweep order.
void GpuSolver::wide_sweep(bool upper, int k, const void *src, void *dst, int ld) {
    const SweepDev &s = m_sweep[upper ? 1 : 0];
    if (s.wide_levels.empty()) return;
    char name[64];
    snprintf(name, sizeof name, "chb_wide_%s_%s", m_elem == 4 ? "f32" : "f64", upper ? "upper" : "lower");
    CUfunction fn = kernel(m_dev, name);
    if (!fn) throw std::runtime_error(driver_error_text());
    const size_t smem = 2 * (size_t) page_bytes_for(m_elem);
    const int block = 32 * sweep_warps(k);
    cuda_check(cuFuncSetAttribute(fn, kFuncAttrMaxDynamicSharedSize, (int) smem));
    int per_sm = max_resident_blocks(m_dev, fn, block, smem) / std::max(1, m_dev->sm_count);
    if (const char *cap = std::getenv("CHB_WIDE_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, std::atoi(cap)));
    for (const std::pair<int, int> &range : s.wide_levels) {
        auto fill = [&](auto &a) {
            using A = std::remove_reference_t<decltype(a)>;
            using Float = std::remove_pointer_t<decltype(A::tmp)>;
            a.n_tasks = range.second;
            a.k = k;
            a.epoch = m_epoch;
            a.src_vec = a.dst_vec = 0;
            a.src_ld = ld;
            a.dst_ld = ld;
            a.slab = (const unsigned char *) s.slab;
            a.tasks = (const SweepTask *) s.wide_tasks + range.first;
            a.done = (int *) s.done;
            a.tmp = (Float *) m_tmp;
            a.src = (const Float *) src;
            a.dst = (Float *) dst;
            a.scratch = (Float *) m_scratch;
            a.rbuf = (Float *) m_rbuf;
            a.cbuf = (Float *) m_cbuf;
        };
        WideArgs<float> a32;
        WideArgs<double> a64;
        void *params[1];
        if (m_elem == 4) {
            fill(a32);
            params[0] = &a32;
        } else {
            fill(a64);
            params[0] = &a64;
        }
        const int grid = std::max(1, std::min(range.second, per_sm * m_dev->sm_count));
        cuda_check(cuLaunchKernel(fn, (unsigned) grid, 1, 1, (unsigned) block, 1, 1, (unsigned) smem, nullptr, params, nullptr));
    }
}

void GpuSolver::sweep(bool upper, int k, const void *src, void *dst, int ld) {
    if (!m_sweep[upper ? 1 : 0].split) split_tasks(m_sweep[upper ? 1 : 0], nullptr);  // a replica's first solve
    // Forward: local tasks (they depend on nothing), the wide levels next to the leaves, then the dependency-driven
    // rest in the persistent kernel.  Backward: the mirror image (nothing depends on the local tasks).
    if (!upper) {
        local_sweep(false, k, src, dst, ld);
        wide_sweep(false, k, src, dst, ld);
    }
    persistent_sweep(upper, k, src, dst, ld);
    if (upper) {
        wide_sweep(true, k, src, dst, ld);
        local_sweep(true, k, src, dst, ld);
    }
}

void GpuSolver::persistent_sweep(bool upper, int k, const void *src, void *dst, int ld) {
    const bool local = true;  // the persistent kernel walks its own table (split_tasks)
    if (m_sweep[upper ? 1 : 0].n_main == 0) return;
    char name[64];
    snprintf(name, sizeof name, "chb_sweep_%s_%s%s%s", m_elem == 4 ? "f32" : "f64", upper ? "upper" : "lower", m_exact ? "_exact" : "",
             m_sweep[upper ? 1 : 0].trace ? "_trace" : "");
    CUfunction fn = kernel(m_dev, name);
    if (!fn) throw std::runtime_error(driver_error_text());
    // the row ring holds the largest local task this kernel still executes (none, usually, with the local-task kernel)
    const int local_rows = local ? std::max(m_sweep[0].main_local_rows, m_sweep[1].main_local_rows)
                                 : std::max(m_sweep[0].max_local_rows, m_sweep[1].max_local_rows);
    const size_t smem = sweep_smem_bytes(m_elem, k, local_rows);
    const int block = 32 * (sweep_warps(k) + 1);
    cuda_check(cuFuncSetAttribute(fn, kFuncAttrMaxDynamicSharedSize, (int) smem));
    cuda_check(cuFuncSetAttribute(fn, kFuncAttrSharedCarveout, 100));  // all of the SM's L1/shared array as shared memory

    int *counters = (int *) m_counters;
    int *ticket = counters + (m_sweeps & 1), *ticket_next = counters + ((m_sweeps + 1) & 1);
    m_sweeps += 1;
    const SweepDev &s = m_sweep[upper ? 1 : 0];

    auto fill = [&](auto &a) {
        using A = std::remove_reference_t<decltype(a)>;
        using Float = std::remove_pointer_t<decltype(A::tmp)>;
        a.n_tasks = local ? s.n_main : s.n_tasks;
        a.k = k;
        a.epoch = m_epoch;
        a.local_rows = local_rows;
        a.ring_rows = sweep_ring_rows(m_elem, k, local_rows);
        a.src_tma = (((uintptr_t) src & 15) == 0 && ((size_t) k * (size_t) m_elem) % 16 == 0 && ((size_t) ld * (size_t) m_elem) % 16 == 0) ? 1 : 0;
        a.src_ld = ld;
        a.dst_ld = ld;
        a.slab = (const unsigned char *) s.slab;
        a.tasks = (const SweepTask *) (local ? s.main_tasks : s.tasks);
        a.done = (int *) s.done;
        a.ticket = ticket;
        a.ticket_next = ticket_next;
        a.tmp = (Float *) m_tmp;
        a.src = (const Float *) src;
        a.dst = (Float *) dst;
        a.scratch = (Float *) m_scratch;
        a.rbuf = (Float *) m_rbuf;
        a.cbuf = (Float *) m_cbuf;
        a.trace = (unsigned long long *) s.trace;
    };
    SweepArgs<float> a32;
    SweepArgs<double> a64;
    void *params[1];
    if (m_elem == 4) {
        fill(a32);
        params[0] = &a32;
    } else {
        fill(a64);
        params[0] = &a64;
    }
    int grid = std::max(1, std::min(local ? s.n_main : s.n_tasks, max_resident_blocks(m_dev, fn, block, smem)));
    if (const char *cap = std::getenv("CHB_SWEEP_CTAS_PER_SM")) grid = std::max(1, std::min(grid, std::atoi(cap) * m_dev->sm_count));
    m_last_grid = grid;
    cuda_check(cuLaunchCooperativeKernel(fn, grid, 1, 1, block, 1, 1, (unsigned) smem, nullptr, params));
}

void GpuSolver::front_sweep(bool upper, int k, const void *src, void *dst, int ld) {
    const SweepDev &s = m_sweep[upper ? 1 : 0];
    char name[64];
    // the tracing kernels (CHB_TRACE=1) also carry the experiment switches of CHB_FRONT_DEBUG
    snprintf(name, sizeof name, "chb_front_%s_c%d_w%d%s", m_elem == 4 ? "f32" : "f64", front_cpl(k), s.n_warps, s.trace ? "_trace" : "");
    CUfunction fn = kernel(m_dev, name);
    if (!fn) throw std::runtime_error(driver_error_text());
    const int max_stage = std::max(m_sweep[0].max_stage_rows, m_sweep[1].max_stage_rows);
    const int ring_rows = front_ring_rows(m_elem, k, s.page_bytes, s.n_warps, max_stage);
    const size_t smem = front_smem_bytes(m_elem, k, s.page_bytes, ring_rows);
    const int block = 32 * (s.n_warps + 1);
    cuda_check(cuFuncSetAttribute(fn, kFuncAttrMaxDynamicSharedSize, (int) smem));
    cuda_check(cuFuncSetAttribute(fn, kFuncAttrSharedCarveout, 100));

    int *counters = (int *) m_counters;
    int *ticket = counters + (m_sweeps & 1), *ticket_next = counters + ((m_sweeps + 1) & 1);
    m_sweeps += 1;
    auto fill = [&](auto &a) {
        using A = std::remove_reference_t<decltype(a)>;
        using Float = std::remove_pointer_t<decltype(A::tmp)>;
        a.n_tasks = s.n_tasks;
        a.k = k;
        a.epoch = m_epoch;
        a.ring_rows = ring_rows;
        a.page_bytes = s.page_bytes;
        a.src_tma = (((uintptr_t) src & 15) == 0 && ((size_t) k * (size_t) m_elem) % 16 == 0 && ((size_t) ld * (size_t) m_elem) % 16 == 0) ? 1 : 0;
        a.dst_vec = (((uintptr_t) dst & 15) == 0 && k % 4 == 0 && ((size_t) ld * (size_t) m_elem) % 16 == 0) ? 1 : 0;
        a.src_ld = ld;
        a.dst_ld = ld;
        a.debug = 0;
        if (const char *v = std::getenv("CHB_FRONT_DEBUG")) a.debug = std::atoi(v);
        a.slab = (const unsigned char *) s.slab;
        a.tasks = (const SweepTask *) s.tasks;
        a.done = (int *) s.done;
        a.ticket = ticket;
        a.ticket_next = ticket_next;
        a.error = counters + 4;
        a.tmp = (Float *) m_tmp;
        a.wbuf = (Float *) m_wbuf;
        a.src = (const Float *) src;
        a.dst = (Float *) dst;
        a.trace = (unsigned long long *) s.trace;
    };
    FrontArgs<float> a32;
    FrontArgs<double> a64;
    void *params[1];
    if (m_elem == 4) {
        fill(a32);
        params[0] = &a32;
    } else {
        fill(a64);
        params[0] = &a64;
    }
    int grid = std::max(1, std::min(s.n_tasks, max_resident_blocks(m_dev, fn, block, smem)));
    if (const char *cap = std::getenv("CHB_SWEEP_CTAS_PER_SM")) grid = std::max(1, std::min(grid, std::atoi(cap) * m_dev->sm_count));
    m_last_grid = grid;
    cuda_check(cuLaunchCooperativeKernel(fn, grid, 1, 1, block, 1, 1, (unsigned) smem, nullptr, params));
}

int GpuSolver::device_error() {
    ContextGuard guard(m_dev);
    int word = 0;
    d2h(&word, (int *) m_counters + 4, sizeof(int));
    sync0();
    return word;
}

void GpuSolver::solve(int k, const void *b, void *x, int ld_b, int ld_x) {
    if (ld_b <= 0) ld_b = k;
    if (ld_x <= 0) ld_x = k;
    if (ld_b < k || ld_x < k) throw std::invalid_argument("row stride smaller than the number of RHS columns");
    if (k != m_k) {
        if (k > 128) throw std::invalid_argument("The number of RHS should be less than 128.");
        if (k < 1) throw std::invalid_argument("The number of RHS should be at least 1.");
    }
    ContextGuard guard(m_dev);
    if (k != m_k) {
        if (padded_rhs(k) > m_kp_cap) {  // work buffers grow, never shrink (the kernels address them with round4(k))
            dev_free(m_tmp);
            dev_free(m_scratch);
            dev_free(m_rbuf);
            dev_free(m_cbuf);
            dev_free(m_wbuf);
            m_kp_cap = 0;  // (stays 0 if an allocation below throws: the next call starts over)
            m_k = 0;
            const int kp_new = padded_rhs(k);
            m_tmp = dev_alloc((size_t) m_elem * (size_t) m_n * (size_t) kp_new);
            if (m_front) {
                const size_t w_rows = (size_t) std::max(m_sweep[0].n_w_rows, m_sweep[1].n_w_rows);
                m_wbuf = dev_alloc((size_t) m_elem * w_rows * (size_t) kp_new);
            } else {
                m_cbuf = dev_alloc((size_t) m_elem * (size_t) m_sweep[0].n_contrib_rows * (size_t) kp_new);
                m_rbuf = dev_alloc((size_t) m_elem * (size_t) m_n * (size_t) kp_new);
                const size_t slots = (size_t) std::max(m_sweep[0].n_slots, m_sweep[1].n_slots);
                m_scratch = dev_alloc((size_t) m_elem * slots * (size_t) kPanelWidth * (size_t) kp_new);
            }
            m_kp_cap = kp_new;
        }
        m_k = k;
    }
    m_epoch += 1;
    if (m_front) {
        front_sweep(false, k, b, nullptr, ld_b);  // tmp = L^-1 P b
        front_sweep(true, k, nullptr, x, ld_x);   // x = P' L^-T tmp
    } else {
        sweep(false, k, b, nullptr, ld_b);
        sweep(true, k, nullptr, x, ld_x);
    }
}

// Host buffers end to end.  RHS columns are independent, so for k >= 64 the columns are processed in two (up to
// four) blocks: the strided H2D copy of block c + 1 (copy stream), the two sweeps of block c (legacy stream) and the
// D2H copy of block c - 1 (second copy stream) overlap; events order them.  PCIe then runs in both directions
// while the sweeps execute.
bool GpuSolver::wide_rhs_allowed() {
    const char *v = std::getenv("CHB_ALLOW_WIDE_RHS");
    return v && std::atoi(v) != 0;
}

void GpuSolver::solve_wide(int k, const void *b, void *x, int ld_b, int ld_x) {
    if (ld_b <= 0) ld_b = k;
    if (ld_x <= 0) ld_x = k;
    if (k <= 128) return solve(k, b, x, ld_b, ld_x);
    if (!wide_rhs_allowed()) throw std::invalid_argument("The number of RHS should be less than 128.");
    for (int c0 = 0; c0 < k; c0 += 128) {  // panels of 128 columns: 512-byte (f32) column offsets, the sweeps' full width
        const int kc = std::min(128, k - c0);
        solve(kc, (const char *) b + (size_t) m_elem * (size_t) c0, (char *) x + (size_t) m_elem * (size_t) c0, ld_b, ld_x);
    }
}

void GpuSolver::solve_host(int k, const void *b_host, void *x_host) {
    if (k > 128 && !wide_rhs_allowed()) throw std::invalid_argument("The number of RHS should be less than 128.");
    if (k < 1) throw std::invalid_argument("The number of RHS should be at least 1.");
    const size_t es = (size_t) m_elem;
    const size_t bytes = es * (size_t) m_n * (size_t) k;
    int n_blocks = k >= 64 ? 2 : 1;  // measured on cfg4: 2 blocks of 64 columns beat 4 of 32 (a sweep is latency-bound, not k-bound)
    if (const char *v = std::getenv("CHB_HOST_BLOCKS")) n_blocks = std::max(1, std::min(std::min(4, k), std::atoi(v)));
    n_blocks = std::max(n_blocks, (k + 127) / 128);  // a block is one solve: at most 128 columns
    const int kb = ((k + n_blocks - 1) / n_blocks + 3) & ~3;
    {
        ContextGuard guard(m_dev);
        if (bytes > m_stage_bytes) {
            dev_free(m_stage_b);
            dev_free(m_stage_x);
            m_stage_bytes = 0;
            m_stage_b = dev_alloc(bytes);
            m_stage_x = dev_alloc(bytes);
            m_stage_bytes = bytes;
        }
        if (n_blocks == 1) h2d(m_stage_b, b_host, bytes);
        if (n_blocks > 1 && !m_copy_in) {
            cuda_check(cuStreamCreate(&m_copy_in, 1));   // CU_STREAM_NON_BLOCKING: no implicit sync with the legacy stream
            cuda_check(cuStreamCreate(&m_copy_out, 1));
        }
        while ((int) m_ev_in.size() < n_blocks) {
            CUevent a = nullptr, c = nullptr;
            cuda_check(cuEventCreate(&a, 2));   // CU_EVENT_DISABLE_TIMING
            cuda_check(cuEventCreate(&c, 2));
            m_ev_in.push_back(a);
            m_ev_out.push_back(c);
        }
    }
    if (n_blocks == 1) {
        solve(k, m_stage_b, m_stage_x);
        ContextGuard guard(m_dev);
        d2h(x_host, m_stage_x, bytes);
        sync0();
        return;
    }
    // column block c: columns [c0, c0 + kc) of the row-major (n, k) host arrays <-> a packed (n, kc) device block
    auto copy_block = [&](bool to_device, int c0, int kc, void *dev, CUstream stream) {
        CUDA_MEMCPY2D d;
        std::memset(&d, 0, sizeof d);
        d.WidthInBytes = es * (size_t) kc;
        d.Height = (size_t) m_n;
        if (to_device) {
            d.srcMemoryType = 1, d.srcHost = b_host, d.srcPitch = es * (size_t) k, d.srcXInBytes = es * (size_t) c0;
            d.dstMemoryType = 2, d.dstDevice = dev, d.dstPitch = es * (size_t) kc;
        } else {
            d.srcMemoryType = 2, d.srcDevice = dev, d.srcPitch = es * (size_t) kc;
            d.dstMemoryType = 1, d.dstHost = x_host, d.dstPitch = es * (size_t) k, d.dstXInBytes = es * (size_t) c0;
        }
        cuda_check(cuMemcpy2DAsync(&d, stream));
    };
    std::vector<int> c0s, kcs;
    for (int c0 = 0; c0 < k; c0 += kb) c0s.push_back(c0), kcs.push_back(std::min(kb, k - c0));
    {
        ContextGuard guard(m_dev);
        for (size_t c = 0; c < c0s.size(); ++c) {
            void *dev_b = (char *) m_stage_b + es * (size_t) m_n * (size_t) c0s[c];
            copy_block(true, c0s[c], kcs[c], dev_b, m_copy_in);
            cuda_check(cuEventRecord(m_ev_in[c], m_copy_in));
        }
    }
    for (size_t c = 0; c < c0s.size(); ++c) {
        void *dev_b = (char *) m_stage_b + es * (size_t) m_n * (size_t) c0s[c];
        void *dev_x = (char *) m_stage_x + es * (size_t) m_n * (size_t) c0s[c];
        {
            ContextGuard guard(m_dev);
            cuda_check(cuStreamWaitEvent(nullptr, m_ev_in[c], 0));
        }
        solve(kcs[c], dev_b, dev_x);
        ContextGuard guard(m_dev);
        cuda_check(cuEventRecord(m_ev_out[c], nullptr));
        cuda_check(cuStreamWaitEvent(m_copy_out, m_ev_out[c], 0));
        copy_block(false, c0s[c], kcs[c], dev_x, m_copy_out);
    }
    ContextGuard guard(m_dev);
    cuda_check(cuStreamSynchronize(m_copy_out));
}

}  // namespace chb
