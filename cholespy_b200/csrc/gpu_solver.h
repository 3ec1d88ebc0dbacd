// Device-resident state and dispatch of the solve path on one B200.
//
// Replaces the GPU branch of the reference solver class
// (reference: src/cholesky_solver.cpp:245-288 upload, :299-376 analyze_cuda, :378-432
// launch_kernel/solve_cuda, members src/cholesky_solver.h:72-98).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "cuda_driver.h"
#include "factor_host.h"

namespace chb {

// One triangular CSR operand in HBM plus its level schedule.
struct TriangleDev {
    void *rowptr = nullptr;  // int32[n+1]
    void *colidx = nullptr;  // int32[Z]
    void *vals = nullptr;    // Float[Z]
    void *order = nullptr;   // int32[n]: row ids grouped by level, ascending inside a level
    int n_levels = 0;
    std::vector<int> level_of_row;  // host copy of the analysis result (bit-exact parity target)
    std::vector<int> order_host;
};

struct ImageBuffer {
    const char *name;
    void *dptr;
    size_t bytes;
};

class GpuSolver {
public:
    // Uploads L (both orientations, values cast to `elem_size` bytes on the host like the
    // reference, src/cholesky_solver.cpp:255-271), runs the level analysis of both triangles.
    GpuSolver(int device, int elem_size, const HostFactor &L);
    // Allocates an image of the given shape without contents (to be filled by a broadcast from a
    // peer's image, then adopt_levels()).  Used for multi-GPU replication.
    GpuSolver(int device, int elem_size, int n, int64_t nnz, int n_levels_lower, int n_levels_upper);
    ~GpuSolver();

    GpuSolver(const GpuSolver &) = delete;
    GpuSolver &operator=(const GpuSolver &) = delete;

    // Enqueues forward + backward sweeps on the legacy default stream; b, x: (n, k) row-major
    // device arrays of the solver's element type.  Throws std::invalid_argument for k > 128.
    void solve(int k, const void *b, void *x);
    // Same, host buffers: H2D, solve, D2H, stream sync (end-to-end entry).
    void solve_host(int k, const void *b_host, void *x_host);

    int n() const { return m_n; }
    int64_t nnz() const { return m_nnz; }
    int elem_size() const { return m_elem; }
    int device() const { return m_dev->ordinal; }
    const TriangleDev &triangle(bool lower) const { return lower ? m_lower : m_upper; }
    const std::vector<int> &perm_host() const { return m_perm_host; }
    int last_grid() const { return m_last_grid; }
    int launches_per_solve() const { return 2; }

    // Replicable device image: everything a peer needs to solve (factor both ways, schedule, perm).
    std::vector<ImageBuffer> image() const;

private:
    void allocate_image();
    void allocate_scratch();
    void upload_triangle(TriangleDev &t, const std::vector<int> &ptr, const std::vector<int> &idx,
                         const std::vector<double> &vals);
    void analyze(TriangleDev &t, bool lower);
    void sweep(bool upper, int k, const void *src, void *dst);

    DeviceState *m_dev = nullptr;
    int m_elem = 4;
    int m_n = 0;
    int64_t m_nnz = 0;
    int m_k = 0;
    int m_epoch = 0;
    int m_sweeps = 0;
    int m_last_grid = 0;

    void *m_perm = nullptr;
    TriangleDev m_lower, m_upper;
    std::vector<int> m_perm_host;

    void *m_ready = nullptr;    // int32[n] epoch flags
    void *m_counters = nullptr; // int32[8]: [0,1] sweep tickets, [2] analysis ticket, [3] analysis max level
    void *m_tmp = nullptr;      // Float[n*k]
    void *m_stage_b = nullptr, *m_stage_x = nullptr;  // device staging for solve_host
    size_t m_stage_bytes = 0;
};

}  // namespace chb
