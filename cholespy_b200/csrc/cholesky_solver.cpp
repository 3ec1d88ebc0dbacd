// See cholesky_solver.h.  Host flow of the constructor: intake (factor_host.cpp) -> CHOLMOD
// (kept as the reference uses it) -> GPU upload + analysis (gpu_solver.cpp).
#include "cholesky_solver.h"

#include <cstdlib>
#include <stdexcept>

#include "disk_cache.h"
#include "factor_host.h"
#include "gpu_solver.h"

namespace chb {
static int g_default_device = -1;
int default_device() {
    if (g_default_device < 0) {
        const char *env = std::getenv("CHOLESPY_B200_DEVICE");
        g_default_device = env ? std::atoi(env) : 0;
    }
    return g_default_device;
}
void set_default_device(int ordinal) { g_default_device = ordinal; }
}  // namespace chb

template <typename Float>
void CholeskySolver<Float>::build(int n_rows, int nnz, const int *ii, const int *jj, const double *x, MatrixType type,
                                  int device) {
    chb::CscMatrix A = chb::intake_matrix(n_rows, nnz, ii, jj, x, (chb::MatrixFormat) type);
    if (m_cpu) {
        chb::factorize_spd(A, nullptr, &m_cholmod);
    } else {
        // CHB_CACHE_DIR set: the factor of a matrix seen before is read back instead of re-running CHOLMOD
        chb::HostFactor L;
        const bool caching = !chb::cache_dir().empty();
        chb::CacheKey key;
        if (caching) key = chb::matrix_key(A);
        if (!caching || !chb::load_factor(key, &L)) {
            chb::factorize_spd(A, &L, nullptr);
            if (caching) chb::store_factor(key, L);
        }
        m_gpu = new chb::GpuSolver(device, (int) sizeof(Float), L);
    }
}

template <typename Float>
CholeskySolver<Float>::CholeskySolver(int n_rows, int nnz, int *ii, int *jj, double *x, MatrixType type, bool cpu)
    : m_cpu(cpu) {
    build(n_rows, nnz, ii, jj, x, type, cpu ? -1 : chb::default_device());
}

template <typename Float>
CholeskySolver<Float>::CholeskySolver(int n_rows, int nnz, const int *ii, const int *jj, const double *x,
                                      MatrixType type, bool cpu, int device)
    : m_cpu(cpu) {
    build(n_rows, nnz, ii, jj, x, type, device);
}

template <typename Float>
CholeskySolver<Float>::~CholeskySolver() {
    delete m_gpu;
    chb::cholmod_state_free(m_cholmod);
}

template <typename Float>
void CholeskySolver<Float>::solve_cuda(int n_rhs, CUdeviceptr b, CUdeviceptr x) {
    if (!m_gpu) throw std::invalid_argument("Input device is CUDA but the solver was initialized for CPU.");
    m_gpu->solve_wide(n_rhs, b, x);
}

template <typename Float>
void CholeskySolver<Float>::solve_cpu(int n_rhs, Float *b, Float *x) {
    if (!m_cholmod) throw std::invalid_argument("Input device is CPU but the solver was initialized for CUDA.");
    chb::cholmod_state_solve<Float>(m_cholmod, n_rhs, b, x);
}

template class CholeskySolver<float>;
template class CholeskySolver<double>;
