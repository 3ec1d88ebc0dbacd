// Host-side matrix intake + CHOLMOD factorisation.  See factor_host.h.
#include "factor_host.h"

#include <algorithm>
#include <cstring>
#include <numeric>

#include "cholmod.h"

namespace chb {

// ---------------------------------------------------------------------------------------------
// Matrix intake: any of COO / CSR / CSC -> canonical CSC (sorted, duplicates summed).
// Behavioural spec: reference src/cholesky_solver.cpp:8-149,160-179.  Duplicates are summed in
// input order (stable), which is the well-defined subset of what the reference's unstable sort does.
// ---------------------------------------------------------------------------------------------
CscMatrix intake_matrix(int n, int nnz, const int *ii, const int *jj, const double *x, MatrixFormat fmt) {
    // Step 1: expand to (row, col) per entry without copying values yet.
    std::vector<int> row_of(nnz), col_of(nnz);
    if (fmt == FMT_COO) {
        std::memcpy(row_of.data(), ii, sizeof(int) * (size_t) nnz);
        std::memcpy(col_of.data(), jj, sizeof(int) * (size_t) nnz);
    } else {
        // ii = pointer array (n+1), jj = minor indices; a malformed pointer array must not index out of bounds
        if (n < 0 || ii[0] != 0 || ii[n] != nnz) throw std::invalid_argument("Sparse matrix: invalid pointer array.");
        for (int major = 0; major < n; ++major)
            if (ii[major + 1] < ii[major]) throw std::invalid_argument("Sparse matrix: invalid pointer array.");
        for (int major = 0; major < n; ++major)
            for (int p = ii[major]; p < ii[major + 1]; ++p) {
                if (fmt == FMT_CSC) { col_of[p] = major; row_of[p] = jj[p]; }
                else                { row_of[p] = major; col_of[p] = jj[p]; }
            }
    }
    for (int e = 0; e < nnz; ++e)
        if (row_of[e] < 0 || row_of[e] >= n || col_of[e] < 0 || col_of[e] >= n)
            throw std::invalid_argument("Sparse matrix: index out of range.");

    // Step 2: stable bucket by column.
    std::vector<int> start(n + 1, 0);
    for (int e = 0; e < nnz; ++e) start[col_of[e] + 1]++;
    std::partial_sum(start.begin(), start.end(), start.begin());
    std::vector<int> order(nnz);
    {
        std::vector<int> cursor(start.begin(), start.end() - 1);
        for (int e = 0; e < nnz; ++e) order[cursor[col_of[e]]++] = e;
    }

    // Step 3: per column stable sort by row, then fold equal rows.
    CscMatrix A;
    A.n = n;
    A.colptr.assign(n + 1, 0);
    A.rowidx.reserve(nnz);
    A.vals.reserve(nnz);
    for (int c = 0; c < n; ++c) {
        int *lo = order.data() + start[c], *hi = order.data() + start[c + 1];
        bool sorted = true;
        for (int *p = lo; p + 1 < hi; ++p)
            if (row_of[p[0]] > row_of[p[1]]) { sorted = false; break; }
        if (!sorted)
            std::stable_sort(lo, hi, [&](int a, int b) { return row_of[a] < row_of[b]; });
        for (int *p = lo; p < hi;) {
            int r = row_of[*p];
            double acc = x[*p++];
            while (p < hi && row_of[*p] == r) acc += x[*p++];
            A.rowidx.push_back(r);
            A.vals.push_back(acc);
        }
        A.colptr[c + 1] = (int) A.rowidx.size();
    }
    return A;
}

// ---------------------------------------------------------------------------------------------
// CHOLMOD
// ---------------------------------------------------------------------------------------------
struct CholmodState {
    cholmod_common common;
    cholmod_factor *factor = nullptr;
    cholmod_dense *rhs = nullptr;  // reusable (n x k) workspace
    int rhs_cols = 0;
    int n = 0;
};

void cholmod_state_free(CholmodState *st) {
    if (!st) return;
    if (st->rhs) cholmod_free_dense(&st->rhs, &st->common);
    if (st->factor) cholmod_free_factor(&st->factor, &st->common);
    cholmod_finish(&st->common);
    delete st;
}

void factorize_spd(const CscMatrix &A, HostFactor *out, CholmodState **keep_state) {
    CholmodState *st = new CholmodState();
    st->n = A.n;
    cholmod_common *cm = &st->common;
    cholmod_start(cm);
    // Same four options as the reference (src/cholesky_solver.cpp:206-209).
    cm->supernodal = CHOLMOD_SIMPLICIAL;
    cm->final_ll = 1;
    cm->nmethods = 1;
    cm->method[0].ordering = CHOLMOD_NESDIS;

    const int n = A.n, nnz = A.nnz();
    cholmod_sparse *M = cholmod_allocate_sparse(n, n, nnz, /*sorted*/ 1, /*packed*/ 1, /*stype*/ -1, CHOLMOD_REAL, cm);
    if (!M) { cholmod_state_free(st); throw std::runtime_error("cholmod_allocate_sparse failed"); }
    std::memcpy(M->p, A.colptr.data(), sizeof(int) * (size_t) (n + 1));
    std::memcpy(M->i, A.rowidx.data(), sizeof(int) * (size_t) nnz);
    std::memcpy(M->x, A.vals.data(), sizeof(double) * (size_t) nnz);

    st->factor = cholmod_analyze(M, cm);
    if (st->factor) cholmod_factorize(M, st->factor, cm);
    const int status = cm->status;
    cholmod_free_sparse(&M, cm);
    if (status == CHOLMOD_NOT_POSDEF) {
        cholmod_state_free(st);
        throw std::invalid_argument("Matrix is not positive definite!");
    }
    if (!st->factor || status < 0) {
        cholmod_state_free(st);
        throw std::runtime_error("CHOLMOD factorisation failed (status " + std::to_string(status) + ")");
    }

    if (keep_state) {  // CPU mode: keep the numeric factor for cholmod_solve
        *keep_state = st;
        return;
    }

    // GPU mode: pull L out as packed CSC (diagonal first per column) and drop CHOLMOD.
    out->n = n;
    out->perm.assign((int *) st->factor->Perm, (int *) st->factor->Perm + n);
    cholmod_sparse *Lc = cholmod_factor_to_sparse(st->factor, cm);  // O(1): steals the factor's arrays
    if (!Lc) { cholmod_state_free(st); throw std::runtime_error("cholmod_factor_to_sparse failed"); }
    const int *Lp = (const int *) Lc->p;
    const int64_t lnz = Lp[n];
    out->colptr.assign(Lp, Lp + n + 1);
    out->rowidx.assign((int *) Lc->i, (int *) Lc->i + lnz);
    out->vals.assign((double *) Lc->x, (double *) Lc->x + lnz);
    cholmod_free_sparse(&Lc, cm);
    cholmod_state_free(st);
}

template <typename Float>
void cholmod_state_solve(CholmodState *st, int k, const Float *b, Float *x) {
    cholmod_common *cm = &st->common;
    const int n = st->n;
    if (k != st->rhs_cols || !st->rhs) {
        if (st->rhs) cholmod_free_dense(&st->rhs, cm);
        st->rhs = cholmod_allocate_dense(n, k, n, CHOLMOD_REAL, cm);
        st->rhs_cols = k;
    }
    double *B = (double *) st->rhs->x;  // column-major
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) B[(size_t) j * n + i] = (double) b[(size_t) i * k + j];
    cholmod_dense *X = cholmod_solve(CHOLMOD_A, st->factor, st->rhs, cm);
    if (!X) throw std::runtime_error("cholmod_solve failed");
    const double *S = (const double *) X->x;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) x[(size_t) i * k + j] = (Float) S[(size_t) j * n + i];
    cholmod_free_dense(&X, cm);
}
template void cholmod_state_solve<float>(CholmodState *, int, const float *, float *);
template void cholmod_state_solve<double>(CholmodState *, int, const double *, double *);

void csc_to_csr(int n, const std::vector<int> &colptr, const std::vector<int> &rowidx, const std::vector<double> &vals,
                std::vector<int> &rowptr, std::vector<int> &colidx, std::vector<double> &rvals) {
    const size_t nnz = (size_t) colptr[n];
    rowptr.assign(n + 1, 0);
    for (size_t p = 0; p < nnz; ++p) rowptr[rowidx[p] + 1]++;
    std::partial_sum(rowptr.begin(), rowptr.end(), rowptr.begin());
    colidx.resize(nnz);
    rvals.resize(nnz);
    std::vector<int> cursor(rowptr.begin(), rowptr.end() - 1);
    for (int c = 0; c < n; ++c)  // ascending columns => each row ends up sorted by column
        for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
            int dst = cursor[rowidx[p]]++;
            colidx[dst] = c;
            rvals[dst] = vals[p];
        }
}

}  // namespace chb
