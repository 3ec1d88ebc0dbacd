// Solver class with the reference's public signature, so the reference's unmodified bindings
// (reference: src/main.cpp:16-110) compile and link against this implementation.
// Mirrors reference src/cholesky_solver.h:8-47 (MatrixType values, constructor, solve_cuda,
// solve_cpu, is_cpu); everything behind it is new (pimpl, no CHOLMOD types in the header).
#pragma once
#include "cuda_driver.h"

enum MatrixType { CSC = 0, CSR, COO };

namespace chb {
class GpuSolver;
struct CholmodState;
// Device ordinal used by solvers built through the reference-signature constructor (the reference
// is hard-wired to device 0, src/cuda_driver.cpp:124).  Default 0, or $CHOLESPY_B200_DEVICE.
int default_device();
void set_default_device(int ordinal);
}  // namespace chb

template <typename Float>
class CholeskySolver {
public:
    // ii/jj/x are HOST arrays (COO: rows/cols; CSR/CSC: pointer array / minor indices); they are only
    // read during the call.  cpu == false builds the device-resident solver on default_device().
    CholeskySolver(int n_rows, int nnz, int *ii, int *jj, double *x, MatrixType type, bool cpu);
    // New: explicit device ordinal (one process per GPU launches).
    CholeskySolver(int n_rows, int nnz, const int *ii, const int *jj, const double *x, MatrixType type, bool cpu,
                   int device);
    ~CholeskySolver();
    CholeskySolver(const CholeskySolver &) = delete;
    CholeskySolver &operator=(const CholeskySolver &) = delete;

    // b, x: (n, n_rhs) row-major device arrays.  Asynchronous on the legacy default stream.
    void solve_cuda(int n_rhs, CUdeviceptr b, CUdeviceptr x);
    // b, x: (n, n_rhs) row-major host arrays; CHOLMOD solve (CPU-mode solvers only).
    void solve_cpu(int n_rhs, Float *b, Float *x);
    bool is_cpu() { return m_cpu; }

    chb::GpuSolver *gpu() { return m_gpu; }  // nullptr for CPU-mode solvers

private:
    void build(int n_rows, int nnz, const int *ii, const int *jj, const double *x, MatrixType type, int device);
    bool m_cpu;
    chb::GpuSolver *m_gpu = nullptr;
    chb::CholmodState *m_cholmod = nullptr;
};
