/* cholespy_b200 -- C ABI of the B200-native cholespy solve path (libcholespy_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point
 * names the reference interface it stands in for (paths relative to the reference checkout).
 * The reference's own bindings (src/main.cpp) call a C++ class; csrc/cholesky_solver.h keeps that
 * class signature on top of the same implementation, and INTEGRATION.md shows both bindings.
 *
 * Conventions: functions returning int return CHB_OK (0) or an error code; the message is
 * available from chb_last_error() (thread-local).  CHB_ERR_INVALID corresponds to the reference's
 * std::invalid_argument (Python ValueError) with the same message text.
 * Dense arrays are row-major (n, n_rhs), of the solver's element type.
 * Device work is enqueued on the legacy default stream of the device's primary context and is
 * asynchronous with respect to the host, exactly like the reference (src/cuda_driver.cpp:96-109).
 */
#ifndef CHOLESPY_B200_H
#define CHOLESPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHB_OK 0
#define CHB_ERR_INVALID 1 /* std::invalid_argument in the reference -> ValueError */
#define CHB_ERR_RUNTIME 2 /* anything else (CUDA failure, CHOLMOD failure, missing GPU) */

#define CHB_F32 0 /* CholeskySolverF */
#define CHB_F64 1 /* CholeskySolverD */

#define CHB_CSC 0 /* same values as the reference's MatrixType, src/cholesky_solver.h:8-12 */
#define CHB_CSR 1
#define CHB_COO 2

#define CHB_DEVICE_CPU (-1) /* build a CPU-mode solver (CHOLMOD solve), reference ctor cpu=true */

typedef struct chb_solver chb_solver;
typedef struct chb_factor chb_factor;

const char *chb_last_error(void);
int chb_version(void);

/* Loads libcuda, retains `device`'s primary context and loads the embedded sm_100a cubin.
 * Replaces init_cuda() (src/cuda_driver.cpp:58-136), which is fixed to device 0. */
int chb_init(int device);
/* Releases every retained primary context.  Replaces shutdown_cuda() (src/cuda_driver.cpp:138-149). */
void chb_shutdown(void);
/* Device ordinal owning a device pointer, or -1. */
int chb_device_of_pointer(const void *dptr);
/* Blocks until the legacy stream of `device` is idle. */
int chb_sync(int device);

/* Constructor.  Replaces CholeskySolver<Float>::CholeskySolver (src/cholesky_solver.cpp:151-198,
 * called from src/main.cpp:63,70).  ii/jj/x are HOST arrays and are only read during the call:
 *   COO: ii = row indices, jj = column indices;  CSR/CSC: ii = pointer array (n_rows+1), jj = indices.
 * device >= 0: factorise on the host with CHOLMOD, upload L, run the GPU level analysis.
 * device == CHB_DEVICE_CPU: keep the CHOLMOD factor for chb_solver_solve_cpu. */
int chb_solver_create(chb_solver **out, int dtype, int n_rows, int nnz, const int32_t *ii, const int32_t *jj,
                      const double *x, int matrix_type, int device);
/* Same, with ii/jj/x living on `device` (what src/main.cpp:55-61 stages through malloc'd host copies). */
int chb_solver_create_from_device_arrays(chb_solver **out, int dtype, int n_rows, int nnz_ii, int nnz, const void *ii_dev,
                                         const void *jj_dev, const void *x_dev, int matrix_type, int device);
/* The hot path taken alone: skips CHOLMOD, consumes a factor P A P' = L L' as the reference's GPU
 * branch does (src/cholesky_solver.cpp:245-288): L in CSC (sorted rows, diagonal first per
 * column, int32 indices, double values), perm[i] = original index of permuted row i. */
int chb_solver_create_from_factor(chb_solver **out, int dtype, int n, int64_t nnz_l, const int32_t *l_colptr,
                                  const int32_t *l_rowidx, const double *l_vals, const int32_t *perm, int device);
/* Replaces CholeskySolver<Float>::~CholeskySolver (src/cholesky_solver.cpp:465-486). */
void chb_solver_destroy(chb_solver *s);

/* x = A^-1 b on the GPU: fused permute + forward L sweep, then fused backward L' sweep + un-permute.
 * Replaces CholeskySolver<Float>::solve_cuda (src/cholesky_solver.cpp:415-432).  n_rhs > 128 fails
 * with CHB_ERR_INVALID "The number of RHS should be less than 128." like the reference. */
int chb_solver_solve_cuda(chb_solver *s, int n_rhs, const void *b_dev, void *x_dev);
/* End-to-end variant with HOST buffers: H2D(b), the same two sweeps, D2H(x), stream sync. */
int chb_solver_solve_host(chb_solver *s, int n_rhs, const void *b_host, void *x_host);
/* CHOLMOD solve on the host for CPU-mode solvers.  Replaces solve_cpu (src/cholesky_solver.cpp:434-463). */
int chb_solver_solve_cpu(chb_solver *s, int n_rhs, const void *b, void *x);
/* Multi-GPU inside the solve call (single process; new: the reference is single-device, src/cuda_driver.cpp:124-125).
 * `devices` = ordinals of n_devices ADDITIONAL B200s: the device image is replicated to them once (ncclBroadcast over
 * NVLink, peer copies where libnccl is absent), and every later chb_solver_solve_cuda with enough columns splits them
 * into contiguous blocks, one per device: the solver's own device works in place in b / x, the others get their block
 * by a strided peer copy, and the legacy stream of the solver's device is joined to all of them before the call
 * returns (same asynchronous semantics as the unsharded call).  n_devices = 0 undoes it.  The environment variable
 * CHB_SHARD_GPUS=N does the same for every CUDA-mode solver at construction (N visible devices), so an unchanged
 * cholespy program uses them.  chb_solver_shard_info: devices in use, replication (0 none, 1 nccl, 2 peer copies),
 * bytes replicated per device, microseconds it took. */
int chb_solver_shard(chb_solver *s, int n_devices, const int32_t *devices);
int chb_solver_shard_info(const chb_solver *s, int64_t *out, int cap);
/* Blocks until the device is idle and reports whether the sweep kernels' watchdog fired (a task dependency that
 * never arrived: CHB_ERR_RUNTIME).  New: the reference's kernels spin forever in that case
 * (kernels/kernels.cu:96-98). */
int chb_solver_check_error(chb_solver *s);
/* Replaces CholeskySolver<Float>::is_cpu (src/cholesky_solver.h:47). */
int chb_solver_is_cpu(const chb_solver *s);

/* Introspection: out[0..] = n, nnz(L), n_levels(lower), n_levels(upper), dtype, device,
 * grid size of the last sweep, kernel launches per solve, then per sweep (forward, backward):
 * n_tasks, slab bytes, scratch slots, task levels, stored values; then 1 if the image is a front image
 * (csrc/front_schedule.h), its update rows and the compute warps of its CTA shape.  Returns the number of values written. */
int chb_solver_info(const chb_solver *s, int64_t *out, int cap);
/* Analysis products of analyze_cuda (src/cholesky_solver.cpp:341-370), for bit-exact parity checks:
 * level_of_row[n] (the reference's level_ind) and order[n] (the reference's levels array). */
int chb_solver_get_analysis(const chb_solver *s, int lower, int32_t *level_of_row, int32_t *order);
int chb_solver_get_perm(const chb_solver *s, int32_t *perm);
/* Debug timeline of the last solve's tasks (only when the solver was created with CHB_TRACE=1 in the
 * environment): 8 words per task {start ns, slab resident ns, end ns, cycles spent waiting on flags,
 * last wait returned ns, blocks done ns, triangle done ns, unused}.
 * Returns the number of words copied. */
int64_t chb_solver_get_trace(chb_solver *s, int upper, uint64_t *out, int64_t cap);

/* Host factorisation only (CHOLMOD, same options as src/cholesky_solver.cpp:204-238), for factor
 * caching and for handing the same factor to several GPUs. */
int chb_factorize(chb_factor **out, int n_rows, int nnz, const int32_t *ii, const int32_t *jj, const double *x,
                  int matrix_type);
int chb_factor_n(const chb_factor *f);
int64_t chb_factor_nnz(const chb_factor *f);
int chb_factor_export(const chb_factor *f, int32_t *perm, int32_t *l_colptr, int32_t *l_rowidx, double *l_vals);
void chb_factor_destroy(chb_factor *f);

/* Multi-GPU replication of the device image (both sweep images: record slabs, task tables, ticket
 * orders): build the solver on one GPU, read its shape, create empty replicas of the same shape on
 * the others, and broadcast each image buffer (NCCL) from the source into the replica.  New; the
 * reference is single-GPU.  shape = CHB_IMAGE_SHAPE_WORDS int64 words. */
#define CHB_IMAGE_SHAPE_WORDS 24
int chb_solver_image_shape(const chb_solver *s, int64_t *shape, int cap);
int chb_solver_create_replica(chb_solver **out, int dtype, const int64_t *shape, int n_words, int device);
int chb_solver_image_count(const chb_solver *s);
int chb_solver_image_buffer(const chb_solver *s, int index, const char **name, void **dptr, uint64_t *bytes);

/* Host-side export of the sweep schedule the kernels stream (csrc/panel_schedule.h): the factor re-laid
 * out as panels / blocks / tasks with a level-ordered ticket list.  Needs no GPU; used by the CPU
 * tests (which interpret it with oracle/panel_emulator.cpp) and for inspection.  New: the reference
 * uploads plain CSR/CSC (src/cholesky_solver.cpp:299-376).
 * exact_order: -1 = library default / environment, 0 = parallel partial sums, 1 = chained (reference order). */
typedef struct chb_schedule chb_schedule;
int chb_schedule_create(chb_schedule **out, int dtype, int n, int64_t nnz_l, const int32_t *l_colptr,
                        const int32_t *l_rowidx, const double *l_vals, const int32_t *perm, int exact_order);
/* info[0..9] = n_tasks, slab bytes, page bytes, n_slots, n_levels, n_values, n_stored_values, n_blocks,
 * n_panels, max_local_rows (largest row count a local task holds in shared memory).  tasks = array of {uint64 offset, uint32 bytes, uint32 level}. */
int chb_schedule_sweep(const chb_schedule *s, int upper, const void **slab, const void **tasks, const int32_t **order,
                       int64_t *info, int cap);
void chb_schedule_destroy(chb_schedule *s);

/* Host-side export of the FRONT schedule (csrc/front_schedule.h), the image the default sweep kernels stream:
 * maximal supernodes as dense block columns, complete subtrees as shared-memory local tasks, update rows
 * instead of per-block gathers.  Needs no GPU; interpreted by oracle/front_emulator.cpp in the CPU tests.
 * info[0..16] = n_tasks, slab bytes, page bytes, tile rows, update rows (W), largest staged record, task levels,
 * n_values, n_stored_values, local tasks, dense tasks, MMA records, assembly entries, supernodes, local task
 * count, ticket-order blocks, compute warps of the CTA shape. */
typedef struct chb_front_schedule chb_front_schedule;
int chb_front_schedule_create(chb_front_schedule **out, int dtype, int n, int64_t nnz_l, const int32_t *l_colptr,
                              const int32_t *l_rowidx, const double *l_vals, const int32_t *perm);
int chb_front_schedule_sweep(const chb_front_schedule *s, int upper, const void **slab, const void **tasks,
                             const int32_t **order, int64_t *info, int cap);
void chb_front_schedule_destroy(chb_front_schedule *s);

#ifdef __cplusplus
}
#endif
#endif /* CHOLESPY_B200_H */
