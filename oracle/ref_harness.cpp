// C-callable harness around the REFERENCE's own solver class, compiled unmodified from the
// reference checkout (see Makefile).  TEST INFRASTRUCTURE ONLY: this is the "reference itself run
// here" used to pin the oracle and, on the GPU box, as the reference GPU/CPU arm.
//
// It does what the reference's bindings do around the class (src/main.cpp:47-72, 98-104): init_cuda,
// push the context, construct with host arrays, solve with raw pointers.
#define private public  // the harness reads the reference's device-side analysis products
#include "cholesky_solver.h"
#undef private

#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <stdexcept>
#include <string>

static std::string g_err;

// The reference does dlopen("libcuda.so") (src/cuda_driver.cpp:68); container images often ship only
// libcuda.so.1.  Its translation units are compiled with -Ddlopen=chb_ref_dlopen (see Makefile).
extern "C" void *chb_ref_dlopen(const char *name, int flags) {
    void *h = dlopen(name, flags);
    if (!h && name && std::strcmp(name, "libcuda.so") == 0) h = dlopen("libcuda.so.1", flags);
    return h;
}

struct RefSolver {
    CholeskySolver<float> *f = nullptr;
    CholeskySolver<double> *d = nullptr;
    int n = 0;
};

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// Returns 1 if the reference's CUDA backend initialised (libcuda present, PTX image loaded).
int ref_init_cuda() {
    if (!init_cuda()) return 0;
    return solve_lower_float != nullptr && analysis_lower != nullptr;
}

// Optional: replace the reference's JIT-ed sm_30 PTX entry points by the same kernels compiled for
// sm_100a (oracle/_ref/ref_kernels_sm100a.cubin), e.g. if the driver refuses the old PTX.
int ref_use_cubin(const void *image) {
    scoped_set_context guard(cu_context);
    CUmodule mod = nullptr;
    if (cuModuleLoadData(&mod, image) != CUDA_SUCCESS) return 0;
    int ok = 1;
    ok &= cuModuleGetFunction(&solve_lower_float, mod, "solve_lower_float") == CUDA_SUCCESS;
    ok &= cuModuleGetFunction(&solve_lower_double, mod, "solve_lower_double") == CUDA_SUCCESS;
    ok &= cuModuleGetFunction(&solve_upper_float, mod, "solve_upper_float") == CUDA_SUCCESS;
    ok &= cuModuleGetFunction(&solve_upper_double, mod, "solve_upper_double") == CUDA_SUCCESS;
    ok &= cuModuleGetFunction(&analysis_lower, mod, "analysis_lower") == CUDA_SUCCESS;
    ok &= cuModuleGetFunction(&analysis_upper, mod, "analysis_upper") == CUDA_SUCCESS;
    return ok;
}

void *ref_create(int is_double, int n_rows, int nnz, int *ii, int *jj, double *x, int type, int cpu) {
    try {
        RefSolver *s = new RefSolver();
        s->n = n_rows;
        if (!cpu) {
            if (!init_cuda()) throw std::runtime_error("reference init_cuda() failed");
            scoped_set_context guard(cu_context);
            if (is_double) s->d = new CholeskySolver<double>(n_rows, nnz, ii, jj, x, (MatrixType) type, false);
            else           s->f = new CholeskySolver<float>(n_rows, nnz, ii, jj, x, (MatrixType) type, false);
        } else {
            if (is_double) s->d = new CholeskySolver<double>(n_rows, nnz, ii, jj, x, (MatrixType) type, true);
            else           s->f = new CholeskySolver<float>(n_rows, nnz, ii, jj, x, (MatrixType) type, true);
        }
        return s;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

int ref_solve_cpu(void *h, int k, void *b, void *x) {
    RefSolver *s = (RefSolver *) h;
    try {
        if (s->d) s->d->solve_cpu(k, (double *) b, (double *) x);
        else      s->f->solve_cpu(k, (float *) b, (float *) x);
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

int ref_solve_cuda(void *h, int k, void *b_dev, void *x_dev) {
    RefSolver *s = (RefSolver *) h;
    try {
        scoped_set_context guard(cu_context);
        if (s->d) s->d->solve_cuda(k, b_dev, x_dev);
        else      s->f->solve_cuda(k, b_dev, x_dev);
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

// Copies the reference's device-side `levels` arrays (analyze_cuda's product) and perm to the host.
int ref_get_schedule(void *h, int *lower_levels, int *upper_levels, int *perm) {
    RefSolver *s = (RefSolver *) h;
    scoped_set_context guard(cu_context);
    CUdeviceptr lo = s->d ? s->d->m_lower_levels_d : s->f->m_lower_levels_d;
    CUdeviceptr up = s->d ? s->d->m_upper_levels_d : s->f->m_upper_levels_d;
    CUdeviceptr pm = s->d ? s->d->m_perm_d : s->f->m_perm_d;
    size_t bytes = sizeof(int) * (size_t) s->n;
    int rc = 0;
    rc |= cuMemcpyAsync(lower_levels, lo, bytes, 0);
    rc |= cuMemcpyAsync(upper_levels, up, bytes, 0);
    rc |= cuMemcpyAsync(perm, pm, bytes, 0);
    // cuMemcpyAsync to pageable host memory on the legacy stream returns after the copy is staged;
    // a following synchronous call makes the data final.
    typedef int (*sync_fn)(void);
    static sync_fn ctx_sync = (sync_fn) dlsym(dlopen("libcuda.so.1", RTLD_LAZY), "cuCtxSynchronize");
    if (ctx_sync) rc |= ctx_sync();
    return rc;
}

void ref_destroy(void *h) {
    RefSolver *s = (RefSolver *) h;
    delete s->f;
    delete s->d;
    delete s;
}

}  // extern "C"
