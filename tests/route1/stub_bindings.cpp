// TEST INFRASTRUCTURE: a stand-in for the reference's binding translation unit (src/main.cpp), written against the
// SAME names and call shapes it uses -- placement new of CholeskySolver<Float>(n_rows, nnz, int*, int*, double*,
// MatrixType, bool) (main.cpp:63,70), solve_cpu / solve_cuda / is_cpu (:92-104), init_cuda(), cu_context,
// scoped_set_context, cuda_check and the cuMemcpyAsync function pointer with CUdeviceptr casts (:51-61) -- so that
// "INTEGRATION.md route 1: compile the unmodified main.cpp together with csrc/" is compile- and link-checked without
// nanobind.  It runs the CPU branch (no GPU needed) and checks the solution of a small SPD system.
#include "cholesky_solver.h"
#include "cuda_driver.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <stdexcept>

template <typename Float>
static int run(const char *name) {
    using Class = CholeskySolver<Float>;
    // 1-D Laplacian + I (tridiagonal SPD), COO with both triangles
    const int n = 50;
    int *ii = (int *) malloc(sizeof(int) * 3 * n), *jj = (int *) malloc(sizeof(int) * 3 * n);
    double *xx = (double *) malloc(sizeof(double) * 3 * n);
    int nnz = 0;
    for (int i = 0; i < n; ++i) {
        ii[nnz] = i, jj[nnz] = i, xx[nnz++] = 3.0;
        if (i + 1 < n) {
            ii[nnz] = i, jj[nnz] = i + 1, xx[nnz++] = -1.0;
            ii[nnz] = i + 1, jj[nnz] = i, xx[nnz++] = -1.0;
        }
    }
    alignas(Class) unsigned char storage[sizeof(Class)];
    Class *self = reinterpret_cast<Class *>(storage);
    MatrixType type = COO;
    new (self) Class(n, nnz, ii, jj, xx, type, true);   // the CPU-init line of main.cpp:70
    free(ii), free(jj), free(xx);
    if (!self->is_cpu()) return 1;
    const int k = 3;
    Float *b = (Float *) malloc(sizeof(Float) * n * k), *x = (Float *) malloc(sizeof(Float) * n * k);
    for (int i = 0; i < n * k; ++i) b[i] = (Float) (1 + i % 7), x[i] = 0;
    self->solve_cpu(k, b, x);
    double worst = 0;
    for (int c = 0; c < k; ++c)
        for (int i = 0; i < n; ++i) {
            double r = 3.0 * x[i * k + c] - (i ? x[(i - 1) * k + c] : 0) - (i + 1 < n ? x[(i + 1) * k + c] : 0) - b[i * k + c];
            worst = std::fmax(worst, std::fabs(r));
        }
    bool threw = false;
    try {
        self->solve_cuda(k, (CUdeviceptr) b, (CUdeviceptr) x);   // main.cpp:103 on a CPU-mode solver
    } catch (const std::invalid_argument &) {
        threw = true;
    }
    self->~Class();
    free(b), free(x);
    printf("%s residual %.3e, solve_cuda on a CPU solver raised: %d\n", name, worst, (int) threw);
    return worst < (sizeof(Float) == 4 ? 1e-4 : 1e-10) && threw ? 0 : 1;
}

// The GPU-init lines of main.cpp:51-61 must compile (they are not executed here: no device).
static void gpu_init_shape(int *dev_indices, size_t n) {
    init_cuda();
    scoped_set_context guard(cu_context);
    int *host = (int *) malloc(n * sizeof(int));
    cuda_check(cuMemcpyAsync((CUdeviceptr) host, (CUdeviceptr) dev_indices, n * sizeof(int), 0));
    free(host);
}

int main(int argc, char **) {
    if (argc > 100) gpu_init_shape(nullptr, 0);
    int rc = run<float>("CholeskySolver<float>") | run<double>("CholeskySolver<double>");
    shutdown_cuda();
    return rc;
}
