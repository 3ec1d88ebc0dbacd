// Device-resident state and dispatch of the solve path on one B200.
//
// Replaces the GPU branch of the reference solver class
// (reference: src/cholesky_solver.cpp:245-288 upload, :299-376 analyze_cuda, :378-432
// launch_kernel/solve_cuda, members src/cholesky_solver.h:72-98).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "cuda_driver.h"
#include "factor_host.h"
#include "front_schedule.h"
#include "panel_schedule.h"

namespace chb {

// Level-set analysis products of one triangle (the reference's level_ind / levels arrays).
struct TriangleAnalysis {
    int n_levels = 0;
    std::vector<int> level_of_row;  // bit-exact parity target
    std::vector<int> order_host;    // row ids grouped by level, ascending inside a level
};

// One sweep's image in HBM (see panel_schedule.h).
struct SweepDev {
    void *slab = nullptr;    // record stream
    void *tasks = nullptr;   // SweepTask[n_tasks], in ticket order (bit 31 of `bytes`: a local task)
    // derived from `tasks` (split_tasks): what the persistent kernel and the local-task kernel walk
    void *main_tasks = nullptr, *local_tasks = nullptr, *wide_tasks = nullptr;
    int n_main = 0, n_local = 0, n_wide = 0;
    std::vector<std::pair<int, int>> wide_levels;  // (first, count) into wide_tasks per wide dependency level, in sweep order
    int local_slab_cap = 0;  // largest record stream among the local-kernel tasks, bytes
    int local_rows_cap = 0;  // most rows a local-kernel task holds in shared memory (forward: its unknowns; backward: + external rows)
    int main_local_rows = 0; // largest local task left to the persistent kernel (sizes its row ring), rows
    bool split = false;
    void *done = nullptr;    // int32[n_tasks]: epoch flags (scratch, not part of the image)
    void *trace = nullptr;   // uint64[8 * n_tasks] when CHB_TRACE=1 (debug timeline), else null
    int n_tasks = 0;
    uint64_t slab_bytes = 0;
    int n_slots = 0;
    int n_task_levels = 0;
    int max_local_rows = 0;  // largest local task (rows held in shared memory)
    int64_t n_contrib_rows = 0;  // rows of the contribution buffer this sweep writes (forward only)
    int64_t n_stored_values = 0;
    // front images (front_schedule.h)
    int64_t n_w_rows = 0;    // rows of the update-row buffer this sweep uses
    int max_stage_rows = 0;  // largest staged record (lower bound of the row ring)
    int page_bytes = 0;
    int n_warps = 0;         // compute warps of the CTA shape the image was built for
};

struct ImageBuffer {
    const char *name;
    void *dptr;
    size_t bytes;
};

// Everything a replica needs to allocate an image of the same shape.
struct ImageShape {
    int64_t n = 0, nnz = 0;
    int64_t n_tasks[2] = {0, 0}, slab_bytes[2] = {0, 0}, n_slots[2] = {0, 0}, n_task_levels[2] = {0, 0};
    int64_t n_levels[2] = {0, 0};  // reference level counts (lower, upper)
    int64_t max_local_rows[2] = {0, 0};
    int64_t exact_order = 0;  // the image was built with ScheduleOptions::exact_order (selects the *_exact kernels)
    int64_t n_contrib_rows = 0;  // rows of the forward sweep's contribution buffer
    int64_t front = 0;           // the image is a front image (front_schedule.h; selects the chb_front_* kernels)
    int64_t n_w_rows[2] = {0, 0}, max_stage_rows[2] = {0, 0};
    int64_t n_warps = 0, page_bytes = 0;
    int64_t local_rows_cap[2] = {0, 0};  // selection rule of the local-task kernel (rows a task may hold), per sweep
    int64_t main_local_rows[2] = {0, 0}; // largest local task left to the persistent kernel, per sweep
    static constexpr int kWords = 24;
    void to_words(int64_t *w) const;
    static ImageShape from_words(const int64_t *w);
};

class GpuSolver {
public:
    // Runs the level analysis of both triangles on the device, lays L out as the panel image
    // (values cast to `elem_size` bytes on the host like the reference,
    // src/cholesky_solver.cpp:255-271) and uploads it.
    GpuSolver(int device, int elem_size, const HostFactor &L);
    // Allocates an image of the given shape without contents (to be filled by a broadcast from a
    // peer's image).  Used for multi-GPU replication.
    GpuSolver(int device, int elem_size, const ImageShape &shape);
    ~GpuSolver();

    GpuSolver(const GpuSolver &) = delete;
    GpuSolver &operator=(const GpuSolver &) = delete;

    // Enqueues forward + backward sweeps on the legacy default stream; b, x: (n, k) row-major
    // device arrays of the solver's element type.  Throws std::invalid_argument for k > 128.
    // ld_b / ld_x: elements between consecutive rows of b / x (0: k), for solving a column block of a wider array in place.
    void solve(int k, const void *b, void *x, int ld_b = 0, int ld_x = 0);
    // Any number of columns as panels of <= 128 (opt-in, see wide_rhs_allowed): lifts the reference's cap
    // (src/cholesky_solver.cpp:419-420) without changing the default behaviour.
    void solve_wide(int k, const void *b, void *x, int ld_b = 0, int ld_x = 0);
    // CHB_ALLOW_WIDE_RHS=1 in the environment: more than 128 RHS columns are solved in column panels instead of raising
    // the reference's "The number of RHS should be less than 128."
    static bool wide_rhs_allowed();
    // Same, host buffers: H2D, solve, D2H, stream sync (end-to-end entry).
    void solve_host(int k, const void *b_host, void *x_host);

    int n() const { return m_n; }
    int64_t nnz() const { return m_nnz; }
    int elem_size() const { return m_elem; }
    int device() const { return m_dev->ordinal; }
    const TriangleAnalysis &triangle(bool lower) const { return lower ? m_lower : m_upper; }
    const std::vector<int> &perm_host() const { return m_perm_host; }
    int last_grid() const { return m_last_grid; }
    int launches_per_solve() const {
        int n = 0;
        for (const SweepDev &s : m_sweep) n += (s.n_main > 0 || !s.split ? 1 : 0) + (s.n_local > 0 ? 1 : 0) + (int) s.wide_levels.size();
        return n;
    }
    bool front() const { return m_front; }
    // Blocks until the legacy stream is idle and returns the sweep watchdog's error word (0: fine; non-zero: a
    // dependency never arrived -- a schedule or kernel bug; the results of the last solves are invalid).
    int device_error();
    const SweepDev &sweep_image(bool upper) const { return m_sweep[upper ? 1 : 0]; }
    ImageShape shape() const;
    // Debug timeline of the last solve (CHB_TRACE=1): 8 x n_tasks words, see kernel_args.h.  Returns words copied.
    size_t copy_trace(bool upper, unsigned long long *out, size_t cap);

    // Replicable device image: everything a peer needs to solve.
    std::vector<ImageBuffer> image() const;

private:
    void allocate_sweep(SweepDev &s);
    void analyze(TriangleAnalysis &t, const std::vector<int> &ptr, const std::vector<int> &idx, bool lower);
    void analyze_etree(const HostFactor &L);
    void upload_sweep(SweepDev &d, const SweepImage &img);
    void sweep(bool upper, int k, const void *src, void *dst, int ld);
    void split_tasks(SweepDev &s, const void *host_tasks);
    void persistent_sweep(bool upper, int k, const void *src, void *dst, int ld);
    void local_sweep(bool upper, int k, const void *src, void *dst, int ld);
    void wide_sweep(bool upper, int k, const void *src, void *dst, int ld);
    void front_sweep(bool upper, int k, const void *src, void *dst, int ld);
    void upload_front(SweepDev &d, const FrontImage &img);

    DeviceState *m_dev = nullptr;
    int m_elem = 4;
    int m_n = 0;
    int64_t m_nnz = 0;
    int m_k = 0;
    int m_epoch = 0;
    int m_sweeps = 0;
    int m_last_grid = 0;
    bool m_exact = false;       // exact-order schedule: packed triangles + IEEE division (the *_exact kernels)
    bool m_front = false;       // front image + chb_front_* kernels (CHB_SWEEP=front); default: the panel image + chb_sweep_* kernels

    TriangleAnalysis m_lower, m_upper;
    std::vector<int> m_perm_host;
    SweepDev m_sweep[2];        // [0] forward (L), [1] backward (L')

    void *m_counters = nullptr; // int32[8]: [0,1] sweep tickets, [2] analysis ticket, [3] analysis max level, [4] watchdog error word
    void *m_wbuf = nullptr;     // Float[n_w_rows*round4(k)]: update rows / partial products of the front sweeps
    void *m_tmp = nullptr;      // Float[n*k]
    void *m_rbuf = nullptr;     // Float[n*round4(k)]: residuals of the inverse super-blocks
    void *m_cbuf = nullptr;     // Float[n_contrib_rows*round4(k)]: forward contributions of the local tasks
    void *m_scratch = nullptr;  // Float[max n_slots * 16 * k]: parked partial tiles
    int m_kp_cap = 0;           // padded RHS width the work buffers are allocated for (grow-only)
    // solve_host: device staging + two copy streams so that the H2D copy of column block c + 1, the sweeps of
    // block c and the D2H copy of block c - 1 overlap
    void *m_stage_b = nullptr, *m_stage_x = nullptr;
    size_t m_stage_bytes = 0;
    CUstream m_copy_in = nullptr, m_copy_out = nullptr;
    std::vector<CUevent> m_ev_in, m_ev_out;  // one pair per column block
};

}  // namespace chb
