// C ABI (include/cholespy_b200.h) over the C++ implementation.
#include "cholespy_b200.h"

#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "cholesky_solver.h"
#include "disk_cache.h"
#include "factor_host.h"
#include "front_schedule.h"
#include "gpu_solver.h"
#include "panel_schedule.h"
#include "shard.h"

#define CHB_API extern "C" __attribute__((visibility("default")))

struct chb_solver {
    int dtype = CHB_F32;
    CholeskySolver<float> *f32 = nullptr;    // full-constructor solvers (own the CHOLMOD state in CPU mode)
    CholeskySolver<double> *f64 = nullptr;
    chb::GpuSolver *bare = nullptr;          // from-factor / replica solvers
    chb::ShardGroup *shards = nullptr;       // RHS columns sharded over further devices (chb_solver_shard / CHB_SHARD_GPUS)
    chb::GpuSolver *gpu() const { return bare ? bare : f32 ? f32->gpu() : f64 ? f64->gpu() : nullptr; }
    bool is_cpu() const { return !gpu(); }
};

struct chb_factor {
    chb::HostFactor L;
};

struct chb_schedule {
    chb::PanelStructure structure;
    chb::SweepImage sweep[2];  // [0] forward (lower), [1] backward (upper)
};

struct chb_front_schedule {
    chb::FrontStructure structure;
    chb::FrontImage sweep[2];  // [0] forward (lower), [1] backward (upper)
};

namespace {
thread_local std::string t_error;

template <typename Fn>
int guarded(Fn &&fn) {
    try {
        cuda_clear_error();
        fn();
        if (cuda_first_error() != CUDA_SUCCESS) {
            t_error = chb::driver_error_text();
            return CHB_ERR_RUNTIME;
        }
        return CHB_OK;
    } catch (const std::invalid_argument &e) {
        t_error = e.what();
        return CHB_ERR_INVALID;
    } catch (const std::bad_alloc &) {
        t_error = "out of host memory";
        return CHB_ERR_RUNTIME;
    } catch (const std::exception &e) {
        t_error = e.what();
        return CHB_ERR_RUNTIME;
    }
}

// CHB_SHARD_GPUS=N in the environment: every CUDA-mode solver shards its multi-RHS solves over N visible devices
// (the solver's own device first), so that an unchanged cholespy program uses them.
void shard_from_env(chb_solver *s) {
    const char *v = std::getenv("CHB_SHARD_GPUS");
    chb::GpuSolver *g = s->gpu();
    if (!v || !g) return;
    const int want = std::atoi(v);
    if (want < 2) return;
    int count = 0;
    if (cuDeviceGetCount(&count) != CUDA_SUCCESS) return;
    std::vector<int> others;
    for (int d = 0; d < count && (int) others.size() + 1 < want; ++d)
        if (d != g->device()) others.push_back(d);
    if (!others.empty()) s->shards = new chb::ShardGroup(g, others);
}

void check_dtype(int dtype) {
    if (dtype != CHB_F32 && dtype != CHB_F64) throw std::invalid_argument("dtype must be CHB_F32 or CHB_F64");
}

void fill_factor(chb::HostFactor &L, int n, int64_t nnz_l, const int32_t *l_colptr, const int32_t *l_rowidx,
                 const double *l_vals, const int32_t *perm) {
    if (n < 0 || l_colptr[n] != nnz_l) throw std::invalid_argument("factor: colptr[n] != nnz");
    L.n = n;
    L.perm.assign(perm, perm + n);
    L.colptr.assign(l_colptr, l_colptr + n + 1);
    L.rowidx.assign(l_rowidx, l_rowidx + nnz_l);
    L.vals.assign(l_vals, l_vals + nnz_l);
    for (int c = 0; c < n; ++c) {
        if (L.colptr[c + 1] <= L.colptr[c] || L.rowidx[L.colptr[c]] != c)
            throw std::invalid_argument("factor: every column must start with its diagonal entry");
        for (int e = L.colptr[c] + 1; e < L.colptr[c + 1]; ++e)
            if (L.rowidx[e] <= L.rowidx[e - 1] || L.rowidx[e] >= n)
                throw std::invalid_argument("factor: row indices must be strictly ascending and below n");
    }
}
}  // namespace

CHB_API void chb_solver_destroy(chb_solver *s);
CHB_API const char *chb_last_error(void) { return t_error.c_str(); }
CHB_API int chb_version(void) { return 100; }

CHB_API int chb_init(int device) {
    return guarded([&] {
        if (!chb::device_state(device)) throw std::runtime_error(chb::driver_error_text());
        if (device == 0) init_cuda();
    });
}
CHB_API void chb_shutdown(void) { shutdown_cuda(); }
CHB_API int chb_device_of_pointer(const void *dptr) { return chb::device_of_pointer(dptr); }
CHB_API int chb_sync(int device) {
    return guarded([&] {
        chb::DeviceState *d = chb::device_state(device);
        if (!d) throw std::runtime_error(chb::driver_error_text());
        scoped_set_context guard(d->ctx);
        cuda_check(cuStreamSynchronize(nullptr));
    });
}

CHB_API int chb_solver_create(chb_solver **out, int dtype, int n_rows, int nnz, const int32_t *ii, const int32_t *jj,
                              const double *x, int matrix_type, int device) {
    return guarded([&] {
        check_dtype(dtype);
        if (matrix_type < 0 || matrix_type > 2) throw std::invalid_argument("unknown matrix type");
        const bool cpu = device < 0;
        chb_solver *s = new chb_solver();
        s->dtype = dtype;
        try {
            if (dtype == CHB_F32)
                s->f32 = new CholeskySolver<float>(n_rows, nnz, ii, jj, x, (MatrixType) matrix_type, cpu, device);
            else
                s->f64 = new CholeskySolver<double>(n_rows, nnz, ii, jj, x, (MatrixType) matrix_type, cpu, device);
            shard_from_env(s);
        } catch (...) {
            chb_solver_destroy(s);
            throw;
        }
        *out = s;
    });
}

CHB_API int chb_solver_create_from_device_arrays(chb_solver **out, int dtype, int n_rows, int len_ii, int nnz,
                                                 const void *ii_dev, const void *jj_dev, const void *x_dev,
                                                 int matrix_type, int device) {
    // Stage the user's matrix to the host (CHOLMOD runs there), as src/main.cpp:55-61 does.
    std::vector<int32_t> ii, jj;
    std::vector<double> x;
    int rc = guarded([&] {
        if (len_ii < 0 || nnz < 0) throw std::invalid_argument("Sparse matrix: negative array length.");
        ii.resize((size_t) len_ii);
        jj.resize((size_t) nnz);
        x.resize((size_t) nnz);
        chb::DeviceState *d = chb::device_state(device);
        if (!d) throw std::runtime_error(chb::driver_error_text());
        scoped_set_context guard(d->ctx);
        cuda_check(cuMemcpyAsync(ii.data(), ii_dev, sizeof(int32_t) * ii.size(), nullptr));
        cuda_check(cuMemcpyAsync(jj.data(), jj_dev, sizeof(int32_t) * jj.size(), nullptr));
        cuda_check(cuMemcpyAsync(x.data(), x_dev, sizeof(double) * x.size(), nullptr));
        cuda_check(cuStreamSynchronize(nullptr));
    });
    if (rc != CHB_OK) return rc;
    return chb_solver_create(out, dtype, n_rows, nnz, ii.data(), jj.data(), x.data(), matrix_type, device);
}

CHB_API int chb_solver_create_from_factor(chb_solver **out, int dtype, int n, int64_t nnz_l, const int32_t *l_colptr,
                                          const int32_t *l_rowidx, const double *l_vals, const int32_t *perm,
                                          int device) {
    return guarded([&] {
        check_dtype(dtype);
        if (device < 0) throw std::invalid_argument("chb_solver_create_from_factor needs a CUDA device");
        chb::HostFactor L;
        fill_factor(L, n, nnz_l, l_colptr, l_rowidx, l_vals, perm);
        chb_solver *s = new chb_solver();
        s->dtype = dtype;
        try {
            s->bare = new chb::GpuSolver(device, dtype == CHB_F32 ? 4 : 8, L);
            shard_from_env(s);
        } catch (...) {
            chb_solver_destroy(s);
            throw;
        }
        *out = s;
    });
}

CHB_API int chb_solver_image_shape(const chb_solver *s, int64_t *shape, int cap) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g) throw std::invalid_argument("CPU-mode solvers have no device image.");
        if (cap < chb::ImageShape::kWords) throw std::invalid_argument("shape buffer too small");
        g->shape().to_words(shape);
    });
}

CHB_API int chb_solver_create_replica(chb_solver **out, int dtype, const int64_t *shape, int n_words, int device) {
    return guarded([&] {
        check_dtype(dtype);
        if (n_words < chb::ImageShape::kWords) throw std::invalid_argument("image shape too short");
        chb_solver *s = new chb_solver();
        s->dtype = dtype;
        try {
            s->bare = new chb::GpuSolver(device, dtype == CHB_F32 ? 4 : 8, chb::ImageShape::from_words(shape));
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}

CHB_API void chb_solver_destroy(chb_solver *s) {
    if (!s) return;
    delete s->shards;
    delete s->f32;
    delete s->f64;
    delete s->bare;
    delete s;
}

CHB_API int chb_solver_solve_cuda(chb_solver *s, int n_rhs, const void *b_dev, void *x_dev) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g) throw std::invalid_argument("Input device is CUDA but the solver was initialized for CPU.");
        if (s->shards) s->shards->solve(n_rhs, b_dev, x_dev);
        else g->solve_wide(n_rhs, b_dev, x_dev);
    });
}

CHB_API int chb_solver_shard(chb_solver *s, int n_devices, const int32_t *devices) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g) throw std::invalid_argument("CPU-mode solvers cannot be sharded over GPUs.");
        delete s->shards;
        s->shards = nullptr;
        if (n_devices <= 0) return;
        s->shards = new chb::ShardGroup(g, std::vector<int>(devices, devices + n_devices));
    });
}

CHB_API int chb_solver_shard_info(const chb_solver *s, int64_t *out, int cap) {
    const chb::ShardGroup *sh = s->shards;
    const int64_t v[] = {sh ? sh->n_devices() : 1, sh ? (std::strcmp(sh->replication(), "nccl") == 0 ? 1 : std::strcmp(sh->replication(), "peer") == 0 ? 2 : 0) : 0,
                         sh ? (int64_t) sh->replicated_bytes() : 0, sh ? (int64_t) (sh->replication_seconds() * 1e6) : 0};
    int m = (int) (sizeof(v) / sizeof(v[0]));
    if (cap < m) m = cap;
    std::memcpy(out, v, sizeof(int64_t) * (size_t) m);
    return m;
}

CHB_API int chb_solver_solve_host(chb_solver *s, int n_rhs, const void *b_host, void *x_host) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g) throw std::invalid_argument("chb_solver_solve_host needs a CUDA-mode solver.");
        g->solve_host(n_rhs, b_host, x_host);
    });
}

CHB_API int chb_solver_solve_cpu(chb_solver *s, int n_rhs, const void *b, void *x) {
    return guarded([&] {
        if (!s->is_cpu()) throw std::invalid_argument("Input device is CPU but the solver was initialized for CUDA.");
        if (s->f32) s->f32->solve_cpu(n_rhs, (float *) b, (float *) x);
        else        s->f64->solve_cpu(n_rhs, (double *) b, (double *) x);
    });
}

CHB_API int chb_solver_check_error(chb_solver *s) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g) return;
        const int word = g->device_error();
        if (word != 0)
            throw std::runtime_error("sweep watchdog: a task dependency never arrived (code " + std::to_string(word) +
                                     "); the results of the last solves are invalid");
    });
}

CHB_API int chb_solver_is_cpu(const chb_solver *s) { return s->is_cpu() ? 1 : 0; }

CHB_API int chb_solver_info(const chb_solver *s, int64_t *out, int cap) {
    chb::GpuSolver *g = s->gpu();
    if (!g) return 0;
    const chb::SweepDev &f = g->sweep_image(false), &b = g->sweep_image(true);
    const int64_t v[] = {g->n(), g->nnz(), g->triangle(true).n_levels, g->triangle(false).n_levels, s->dtype,
                         g->device(), g->last_grid(), g->launches_per_solve(),
                         f.n_tasks, (int64_t) f.slab_bytes, f.n_slots, f.n_task_levels, f.n_stored_values,
                         b.n_tasks, (int64_t) b.slab_bytes, b.n_slots, b.n_task_levels, b.n_stored_values,
                         g->front() ? 1 : 0, f.n_w_rows, f.n_warps};
    int m = (int) (sizeof(v) / sizeof(v[0]));
    if (cap < m) m = cap;
    std::memcpy(out, v, sizeof(int64_t) * (size_t) m);
    return m;
}

CHB_API int chb_solver_get_analysis(const chb_solver *s, int lower, int32_t *level_of_row, int32_t *order) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g) throw std::invalid_argument("CPU-mode solvers have no GPU analysis.");
        const chb::TriangleAnalysis &t = g->triangle(lower != 0);
        if (t.level_of_row.empty()) throw std::runtime_error("analysis products are not kept on replicas");
        std::memcpy(level_of_row, t.level_of_row.data(), sizeof(int32_t) * (size_t) g->n());
        std::memcpy(order, t.order_host.data(), sizeof(int32_t) * (size_t) g->n());
    });
}

CHB_API int64_t chb_solver_get_trace(chb_solver *s, int upper, uint64_t *out, int64_t cap) {
    chb::GpuSolver *g = s->gpu();
    if (!g) return 0;
    return (int64_t) g->copy_trace(upper != 0, (unsigned long long *) out, (size_t) cap);
}

CHB_API int chb_solver_get_perm(const chb_solver *s, int32_t *perm) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g || g->perm_host().empty()) throw std::runtime_error("permutation not available on this solver");
        std::memcpy(perm, g->perm_host().data(), sizeof(int32_t) * (size_t) g->n());
    });
}

CHB_API int chb_factorize(chb_factor **out, int n_rows, int nnz, const int32_t *ii, const int32_t *jj, const double *x,
                          int matrix_type) {
    return guarded([&] {
        if (matrix_type < 0 || matrix_type > 2) throw std::invalid_argument("unknown matrix type");
        chb::CscMatrix A = chb::intake_matrix(n_rows, nnz, ii, jj, x, (chb::MatrixFormat) matrix_type);
        chb_factor *f = new chb_factor();
        try {
            // CHB_CACHE_DIR set: the factor of a matrix seen before is read back instead of re-running CHOLMOD
            const bool caching = !chb::cache_dir().empty();
            const chb::CacheKey key = caching ? chb::matrix_key(A) : chb::CacheKey();
            if (!caching || !chb::load_factor(key, &f->L)) {
                chb::factorize_spd(A, &f->L, nullptr);
                if (caching) chb::store_factor(key, f->L);
            }
        } catch (...) {
            delete f;
            throw;
        }
        *out = f;
    });
}
CHB_API int chb_factor_n(const chb_factor *f) { return f->L.n; }
CHB_API int64_t chb_factor_nnz(const chb_factor *f) { return f->L.nnz(); }
CHB_API int chb_factor_export(const chb_factor *f, int32_t *perm, int32_t *l_colptr, int32_t *l_rowidx,
                              double *l_vals) {
    const chb::HostFactor &L = f->L;
    std::memcpy(perm, L.perm.data(), sizeof(int32_t) * (size_t) L.n);
    std::memcpy(l_colptr, L.colptr.data(), sizeof(int32_t) * (size_t) (L.n + 1));
    std::memcpy(l_rowidx, L.rowidx.data(), sizeof(int32_t) * (size_t) L.nnz());
    std::memcpy(l_vals, L.vals.data(), sizeof(double) * (size_t) L.nnz());
    return CHB_OK;
}
CHB_API void chb_factor_destroy(chb_factor *f) { delete f; }

CHB_API int chb_solver_image_count(const chb_solver *s) {
    chb::GpuSolver *g = s->gpu();
    return g ? (int) g->image().size() : 0;
}
CHB_API int chb_solver_image_buffer(const chb_solver *s, int index, const char **name, void **dptr, uint64_t *bytes) {
    return guarded([&] {
        chb::GpuSolver *g = s->gpu();
        if (!g) throw std::invalid_argument("CPU-mode solvers have no device image.");
        std::vector<chb::ImageBuffer> img = g->image();
        if (index < 0 || index >= (int) img.size()) throw std::invalid_argument("image buffer index out of range");
        *name = img[index].name;
        *dptr = img[index].dptr;
        *bytes = img[index].bytes;
    });
}

CHB_API int chb_schedule_create(chb_schedule **out, int dtype, int n, int64_t nnz_l, const int32_t *l_colptr,
                                const int32_t *l_rowidx, const double *l_vals, const int32_t *perm, int exact_order) {
    return guarded([&] {
        check_dtype(dtype);
        chb::HostFactor L;
        fill_factor(L, n, nnz_l, l_colptr, l_rowidx, l_vals, perm);
        chb::ScheduleOptions opt = chb::schedule_options_from_env();
        if (exact_order >= 0) opt.exact_order = exact_order != 0;
        chb_schedule *s = new chb_schedule();
        try {
            chb::build_panel_structure(L, opt, &s->structure);
            for (int upper = 0; upper < 2; ++upper)
                chb::build_sweep_image(L, s->structure, dtype == CHB_F32 ? 4 : 8, upper != 0, opt, &s->sweep[upper]);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}
CHB_API int chb_schedule_sweep(const chb_schedule *s, int upper, const void **slab, const void **tasks,
                               const int32_t **order, int64_t *info, int cap) {
    const chb::SweepImage &img = s->sweep[upper ? 1 : 0];
    *slab = img.slab.data();
    *tasks = img.tasks.data();
    *order = img.order.data();
    const int64_t v[] = {(int64_t) img.tasks.size(), (int64_t) img.slab.size(), img.page_bytes, img.n_slots, img.n_levels,
                         img.n_values, img.n_stored_values, img.n_blocks, s->structure.n_panels, img.max_local_rows};
    int m = (int) (sizeof(v) / sizeof(v[0]));
    if (cap < m) m = cap;
    std::memcpy(info, v, sizeof(int64_t) * (size_t) m);
    return m;
}
CHB_API void chb_schedule_destroy(chb_schedule *s) { delete s; }

CHB_API int chb_front_schedule_create(chb_front_schedule **out, int dtype, int n, int64_t nnz_l, const int32_t *l_colptr,
                                      const int32_t *l_rowidx, const double *l_vals, const int32_t *perm) {
    return guarded([&] {
        check_dtype(dtype);
        chb::HostFactor L;
        fill_factor(L, n, nnz_l, l_colptr, l_rowidx, l_vals, perm);
        const int elem = dtype == CHB_F32 ? 4 : 8;
        const chb::FrontOptions opt = chb::front_options_from_env(elem);
        chb_front_schedule *s = new chb_front_schedule();
        try {
            chb::build_front_structure(L, elem, opt, &s->structure);
            for (int upper = 0; upper < 2; ++upper) chb::build_front_image(L, s->structure, elem, upper != 0, opt, &s->sweep[upper]);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}
CHB_API int chb_front_schedule_sweep(const chb_front_schedule *s, int upper, const void **slab, const void **tasks,
                                     const int32_t **order, int64_t *info, int cap) {
    const chb::FrontImage &img = s->sweep[upper ? 1 : 0];
    *slab = img.slab.data();
    *tasks = img.tasks.data();
    *order = img.order.data();
    const int64_t v[] = {(int64_t) img.tasks.size(), (int64_t) img.slab.size(), img.page_bytes, img.tile_rows, img.n_w_rows,
                         img.max_stage_rows, img.n_levels, img.n_values, img.n_stored_values, img.n_local_tasks,
                         img.n_dense_tasks, img.n_mma, img.n_asm_entries, s->structure.ns, (int64_t) s->structure.locals.size(),
                         s->structure.n_blocks, img.n_warps, img.n_local_levels, img.n_local_rows,
                         img.local_phase_groups[0], img.local_phase_groups[1], img.local_phase_groups[2], img.local_phase_groups[3],
                         img.local_phase_rows[0], img.local_phase_rows[1], img.local_phase_rows[2], img.local_phase_rows[3]};
    int m = (int) (sizeof(v) / sizeof(v[0]));
    if (cap < m) m = cap;
    std::memcpy(info, v, sizeof(int64_t) * (size_t) m);
    return m;
}
CHB_API void chb_front_schedule_destroy(chb_front_schedule *s) { delete s; }
