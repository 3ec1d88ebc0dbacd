/* CPU oracle for the cholespy GPU solve path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of what the reference's device kernels and analysis code
 * compute, used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.
 * Nothing in the product (cholespy_b200/) links, imports or calls this file.
 *
 * Pinned against the reference itself (not "parity unpinned"):
 *   - tests/golden/ holds level arrays, permutations and solutions produced by the reference's own
 *     CHOLMOD build and by this oracle; tests/golden/make_golden.py regenerates them and checks
 *     them against the reference's solve_cpu (oracle/_ref) and scipy, in the build container;
 *   - on the GPU box the parity tests also run the reference's own kernels (oracle/_ref) next to
 *     this oracle.
 *
 * Each function cites the reference lines it follows (paths relative to the reference checkout).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Level of every row of a triangular CSR operand.
 *   descending == 0: follows analysis_lower, kernels/kernels.cu:8-33  (deps: all but the LAST entry)
 *   descending == 1: follows analysis_upper, kernels/kernels.cu:35-60 (deps: all but the FIRST entry,
 *                    rows visited from n-1 down to 0)
 * The GPU kernels resolve rows in ticket order with spin-waits; sequentially that is one pass in the
 * same order.  Returns n_levels = max level + 1 (src/cholesky_solver.cpp:345-346). */
int chb_oracle_levels(int n, const int32_t *rowptr, const int32_t *colidx, int descending, int32_t *level_of_row) {
    int max_level = 0;
    for (int t = 0; t < n; ++t) {
        const int row = descending ? n - 1 - t : t;
        const int start = rowptr[row], end = rowptr[row + 1] - 1; /* inclusive end, as in the kernels */
        int lvl = 0;
        if (!descending) {
            for (int i = start; i < end; ++i) {
                const int dep = level_of_row[colidx[i]];
                if (lvl <= dep) lvl = dep + 1;
            }
        } else {
            for (int i = end; i > start; --i) {
                const int dep = level_of_row[colidx[i]];
                if (lvl <= dep) lvl = dep + 1;
            }
        }
        level_of_row[row] = lvl;
        if (lvl > max_level) max_level = lvl;
    }
    return max_level + 1;
}

/* Rows grouped by level, ascending row index inside a level.
 * Follows the host counting sort in analyze_cuda, src/cholesky_solver.cpp:348-367. */
void chb_oracle_level_order(int n, int n_levels, const int32_t *level_of_row, int32_t *order) {
    int32_t *ptr = (int32_t *) calloc((size_t) n_levels + 1, sizeof(int32_t));
    for (int i = 0; i < n; ++i) ptr[1 + level_of_row[i]]++;
    for (int l = 0; l < n_levels; ++l) ptr[l + 1] += ptr[l];
    for (int i = 0; i < n; ++i) order[ptr[level_of_row[i]]++] = i;
    free(ptr);
}

/* Sweeps.  `use_fma` selects r = fma(-val, x, r) (what ptxas contracts the reference's
 * `r -= val * x`, kernels/kernels.cu:113-114,181-182, into) or a separately rounded multiply and
 * subtract.  This file must be compiled with -ffp-contract=off so that the choice is explicit. */
#define DEFINE_SWEEPS(SUFFIX, T, FMA)                                                                                  \
    /* tmp = L^-1 (P b): follows solve_lower, kernels/kernels.cu:81-120 (diag = last entry of the row,            */ \
    /* b gathered through perm, entries in ascending column order).                                                */ \
    void chb_oracle_lower_##SUFFIX(int n, int k, const int32_t *rowptr, const int32_t *colidx, const T *vals,         \
                                   const int32_t *perm, const T *b, T *tmp, int use_fma) {                            \
        for (int row = 0; row < n; ++row) {                                                                            \
            const int start = rowptr[row], end = rowptr[row + 1] - 1;                                                  \
            const T diag = vals[end];                                                                                  \
            for (int j = 0; j < k; ++j) {                                                                              \
                T r = b[(size_t) perm[row] * k + j];                                                                   \
                for (int i = start; i < end; ++i) {                                                                    \
                    const T x = tmp[(size_t) colidx[i] * k + j];                                                       \
                    r = use_fma ? FMA(-vals[i], x, r) : r - vals[i] * x;                                               \
                }                                                                                                      \
                tmp[(size_t) row * k + j] = r / diag;                                                                  \
            }                                                                                                          \
        }                                                                                                              \
    }                                                                                                                  \
    /* tmp = L^-T tmp, x[perm] = tmp: follows solve_upper, kernels/kernels.cu:149-190 (CSR of L', diag = first      */ \
    /* entry, entries walked from the row end backwards, result scattered through perm).                           */ \
    void chb_oracle_upper_##SUFFIX(int n, int k, const int32_t *rowptr, const int32_t *colidx, const T *vals,         \
                                   const int32_t *perm, T *tmp, T *x_out, int use_fma) {                              \
        for (int row = n - 1; row >= 0; --row) {                                                                       \
            const int start = rowptr[row], end = rowptr[row + 1] - 1;                                                  \
            const T diag = vals[start];                                                                                \
            for (int j = 0; j < k; ++j) {                                                                              \
                T r = tmp[(size_t) row * k + j];                                                                       \
                for (int i = end; i > start; --i) {                                                                    \
                    const T x = tmp[(size_t) colidx[i] * k + j];                                                       \
                    r = use_fma ? FMA(-vals[i], x, r) : r - vals[i] * x;                                               \
                }                                                                                                      \
                r = r / diag;                                                                                          \
                tmp[(size_t) row * k + j] = r;                                                                         \
                x_out[(size_t) perm[row] * k + j] = r;                                                                 \
            }                                                                                                          \
        }                                                                                                              \
    }

DEFINE_SWEEPS(f32, float, fmaf)
DEFINE_SWEEPS(f64, double, fma)

/* CSC of a lower-triangular factor -> CSR (sorted columns, diagonal last), i.e. what
 * cholmod_transpose(lower_csc, 1) hands to analyze_cuda, src/cholesky_solver.cpp:252,277. */
void chb_oracle_csc_to_csr(int n, const int32_t *colptr, const int32_t *rowidx, const double *vals, int32_t *rowptr,
                           int32_t *colidx, double *rvals) {
    const int64_t nnz = colptr[n];
    memset(rowptr, 0, sizeof(int32_t) * ((size_t) n + 1));
    for (int64_t p = 0; p < nnz; ++p) rowptr[rowidx[p] + 1]++;
    for (int i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
    int32_t *cur = (int32_t *) malloc(sizeof(int32_t) * (size_t) n);
    memcpy(cur, rowptr, sizeof(int32_t) * (size_t) n);
    for (int c = 0; c < n; ++c)
        for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
            const int dst = cur[rowidx[p]]++;
            colidx[dst] = c;
            rvals[dst] = vals[p];
        }
    free(cur);
}
