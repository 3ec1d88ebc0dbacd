// Device-resident state and dispatch of the solve path.  See gpu_solver.h.
#include "gpu_solver.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "kernel_args.h"

namespace chb {

namespace {

constexpr int kBlock = 256;  // threads per CTA of the persistent sweeps (8 warps)

void *dev_alloc(size_t bytes) {
    void *p = nullptr;
    cuda_check(cuMemAlloc(&p, std::max<size_t>(bytes, 16)));
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

}  // namespace

void GpuSolver::allocate_image() {
    m_perm = dev_alloc(sizeof(int) * (size_t) m_n);
    for (TriangleDev *t : {&m_lower, &m_upper}) {
        t->rowptr = dev_alloc(sizeof(int) * (size_t) (m_n + 1));
        t->colidx = dev_alloc(sizeof(int) * (size_t) m_nnz);
        t->vals = dev_alloc((size_t) m_elem * (size_t) m_nnz);
        t->order = dev_alloc(sizeof(int) * (size_t) m_n);
    }
}

void GpuSolver::allocate_scratch() {
    m_ready = dev_alloc(sizeof(int) * (size_t) m_n);
    m_counters = dev_alloc(sizeof(int) * 8);
    cuda_check(cuMemsetD32Async(m_ready, 0, (size_t) m_n, nullptr));
    cuda_check(cuMemsetD32Async(m_counters, 0, 8, nullptr));
}

void GpuSolver::upload_triangle(TriangleDev &t, const std::vector<int> &ptr, const std::vector<int> &idx,
                                const std::vector<double> &vals) {
    h2d(t.rowptr, ptr.data(), sizeof(int) * ptr.size());
    h2d(t.colidx, idx.data(), sizeof(int) * idx.size());
    if (m_elem == 8) {
        h2d(t.vals, vals.data(), sizeof(double) * vals.size());
        sync0();
    } else {
        std::vector<float> narrow(vals.size());
        for (size_t i = 0; i < vals.size(); ++i) narrow[i] = (float) vals[i];
        h2d(t.vals, narrow.data(), sizeof(float) * narrow.size());
        sync0();  // `narrow` dies here
    }
}

void GpuSolver::analyze(TriangleDev &t, bool lower) {
    int *counters = (int *) m_counters;
    void *level_d = dev_alloc(sizeof(int) * (size_t) m_n);
    cuda_check(cuMemsetD8Async(level_d, 0xff, sizeof(int) * (size_t) m_n, nullptr));  // all -1
    cuda_check(cuMemsetD32Async(counters + 2, 0, 2, nullptr));

    AnalysisArgs args;
    args.n = m_n;
    args.descending = lower ? 0 : 1;
    args.rowptr = (const int *) t.rowptr;
    args.colidx = (const int *) t.colidx;
    args.level = (int *) level_d;
    args.max_level = counters + 3;
    args.ticket = counters + 2;
    CUfunction fn = kernel(m_dev, "chb_analysis");
    if (!fn) throw std::runtime_error(driver_error_text());
    const int grid = std::max(1, max_resident_blocks(m_dev, fn, kBlock, 0));
    void *params[] = {&args};
    cuda_check(cuLaunchCooperativeKernel(fn, grid, 1, 1, kBlock, 1, 1, 0, nullptr, params));

    t.level_of_row.resize(m_n);
    int max_level = 0;
    d2h(t.level_of_row.data(), level_d, sizeof(int) * (size_t) m_n);
    d2h(&max_level, counters + 3, sizeof(int));
    sync0();
    dev_free(level_d);
    t.n_levels = max_level + 1;
    for (int r = 0; r < m_n; ++r)
        if (t.level_of_row[r] < 0 || t.level_of_row[r] > max_level)
            throw std::runtime_error("level analysis produced an out-of-range level (GPU kernel failure)");
    bucket_rows_by_level(t.level_of_row, t.n_levels, t.order_host);
    h2d(t.order, t.order_host.data(), sizeof(int) * (size_t) m_n);
    sync0();
}

GpuSolver::GpuSolver(int device, int elem_size, const HostFactor &L)
    : m_dev(require_device(device)), m_elem(elem_size), m_n(L.n), m_nnz(L.nnz()) {
    ContextGuard guard(m_dev);
    allocate_image();
    allocate_scratch();
    m_perm_host = L.perm;
    h2d(m_perm, L.perm.data(), sizeof(int) * (size_t) m_n);
    {
        // CSR of L (diagonal last per row): the forward operand.
        std::vector<int> rowptr, colidx;
        std::vector<double> rvals;
        csc_to_csr(L.n, L.colptr, L.rowidx, L.vals, rowptr, colidx, rvals);
        upload_triangle(m_lower, rowptr, colidx, rvals);
    }
    // CSC of L == CSR of L' (diagonal first per row): the backward operand.
    upload_triangle(m_upper, L.colptr, L.rowidx, L.vals);
    analyze(m_lower, true);
    analyze(m_upper, false);
    if (cuda_first_error() != CUDA_SUCCESS) throw std::runtime_error(driver_error_text());
}

GpuSolver::GpuSolver(int device, int elem_size, int n, int64_t nnz, int n_levels_lower, int n_levels_upper)
    : m_dev(require_device(device)), m_elem(elem_size), m_n(n), m_nnz(nnz) {
    ContextGuard guard(m_dev);
    allocate_image();
    allocate_scratch();
    m_lower.n_levels = n_levels_lower;
    m_upper.n_levels = n_levels_upper;
    sync0();
    if (cuda_first_error() != CUDA_SUCCESS) throw std::runtime_error(driver_error_text());
}

GpuSolver::~GpuSolver() {
    ContextGuard guard(m_dev);
    sync0();
    dev_free(m_perm);
    for (TriangleDev *t : {&m_lower, &m_upper}) {
        dev_free(t->rowptr);
        dev_free(t->colidx);
        dev_free(t->vals);
        dev_free(t->order);
    }
    dev_free(m_ready);
    dev_free(m_counters);
    dev_free(m_tmp);
    dev_free(m_stage_b);
    dev_free(m_stage_x);
}

std::vector<ImageBuffer> GpuSolver::image() const {
    const size_t ni = sizeof(int) * (size_t) m_n, zi = sizeof(int) * (size_t) m_nnz;
    const size_t zv = (size_t) m_elem * (size_t) m_nnz;
    return {
        {"perm", m_perm, ni},
        {"lower.rowptr", m_lower.rowptr, ni + sizeof(int)},
        {"lower.colidx", m_lower.colidx, zi},
        {"lower.vals", m_lower.vals, zv},
        {"lower.order", m_lower.order, ni},
        {"upper.rowptr", m_upper.rowptr, ni + sizeof(int)},
        {"upper.colidx", m_upper.colidx, zi},
        {"upper.vals", m_upper.vals, zv},
        {"upper.order", m_upper.order, ni},
    };
}

template <typename Float>
static void fill_args(SolveArgs<Float> &a, int n, int k, bool upper, int epoch, const TriangleDev &t, void *ready,
                      int *ticket, int *ticket_next, void *tmp, const void *src, void *dst, const void *perm) {
    a.n = n;
    a.k = k;
    a.upper = upper ? 1 : 0;
    a.epoch = epoch;
    a.order = (const int *) t.order;
    a.rowptr = (const int *) t.rowptr;
    a.colidx = (const int *) t.colidx;
    a.vals = (const Float *) t.vals;
    a.ready = (int *) ready;
    a.ticket = ticket;
    a.ticket_next = ticket_next;
    a.tmp = (Float *) tmp;
    a.src = (const Float *) src;
    a.dst = (Float *) dst;
    a.perm = (const int *) perm;
}

void GpuSolver::sweep(bool upper, int k, const void *src, void *dst) {
    char name[64];
    const char *ty = m_elem == 4 ? "f32" : "f64";
    if (k <= 8) {
        const int kmax = k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8;
        snprintf(name, sizeof name, "chb_solve_narrow_%s_%d", ty, kmax);
    } else {
        snprintf(name, sizeof name, "chb_solve_wide_%s_%d", ty, (k + 31) / 32);
    }
    CUfunction fn = kernel(m_dev, name);
    if (!fn) throw std::runtime_error(driver_error_text());

    m_epoch += 1;
    int *counters = (int *) m_counters;
    int *ticket = counters + (m_sweeps & 1), *ticket_next = counters + ((m_sweeps + 1) & 1);
    m_sweeps += 1;
    const TriangleDev &t = upper ? m_upper : m_lower;

    SolveArgs<float> a32;
    SolveArgs<double> a64;
    void *params[1];
    if (m_elem == 4) {
        fill_args(a32, m_n, k, upper, m_epoch, t, m_ready, ticket, ticket_next, m_tmp, src, dst, m_perm);
        params[0] = &a32;
    } else {
        fill_args(a64, m_n, k, upper, m_epoch, t, m_ready, ticket, ticket_next, m_tmp, src, dst, m_perm);
        params[0] = &a64;
    }
    const int warps_per_block = kBlock / 32;
    const int needed = (m_n + warps_per_block - 1) / warps_per_block;
    const int grid = std::max(1, std::min(needed, max_resident_blocks(m_dev, fn, kBlock, 0)));
    m_last_grid = grid;
    cuda_check(cuLaunchCooperativeKernel(fn, grid, 1, 1, kBlock, 1, 1, 0, nullptr, params));
}

void GpuSolver::solve(int k, const void *b, void *x) {
    if (k != m_k) {
        if (k > 128) throw std::invalid_argument("The number of RHS should be less than 128.");
        if (k < 1) throw std::invalid_argument("The number of RHS should be at least 1.");
    }
    ContextGuard guard(m_dev);
    if (k != m_k) {
        dev_free(m_tmp);
        m_tmp = dev_alloc((size_t) m_elem * (size_t) m_n * (size_t) k);
        m_k = k;
    }
    sweep(false, k, b, nullptr);  // tmp = L^-1 P b
    sweep(true, k, nullptr, x);   // x = P' L^-T tmp
}

void GpuSolver::solve_host(int k, const void *b_host, void *x_host) {
    const size_t bytes = (size_t) m_elem * (size_t) m_n * (size_t) k;
    {
        ContextGuard guard(m_dev);
        if (bytes > m_stage_bytes) {
            dev_free(m_stage_b);
            dev_free(m_stage_x);
            m_stage_b = dev_alloc(bytes);
            m_stage_x = dev_alloc(bytes);
            m_stage_bytes = bytes;
        }
        h2d(m_stage_b, b_host, bytes);
    }
    solve(k, m_stage_b, m_stage_x);
    ContextGuard guard(m_dev);
    d2h(x_host, m_stage_x, bytes);
    sync0();
}

}  // namespace chb
