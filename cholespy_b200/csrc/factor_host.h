// Host-side matrix intake and CHOLMOD factorisation (not on the hot path).
//
// Mirrors what the reference constructor does before its GPU branch starts
// (reference: src/cholesky_solver.cpp:151-242): COO/CSR -> CSC, sort row indices,
// sum duplicates, then CHOLMOD simplicial LL' with NESDIS ordering.  The implementation
// here is new; only the observable behaviour (same CHOLMOD options, same error text) is kept.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace chb {

enum MatrixFormat { FMT_CSC = 0, FMT_CSR = 1, FMT_COO = 2 };  // same values as reference MatrixType

// Compressed sparse column matrix, sorted row indices, no duplicates.
struct CscMatrix {
    int n = 0;
    std::vector<int> colptr;   // n+1
    std::vector<int> rowidx;   // nnz
    std::vector<double> vals;  // nnz
    int nnz() const { return colptr.empty() ? 0 : colptr[n]; }
};

// Builds a sorted, duplicate-free CSC copy of the user's matrix (never touches the inputs).
CscMatrix intake_matrix(int n_rows, int nnz, const int *ii, const int *jj, const double *x, MatrixFormat fmt);

// The Cholesky factor as the GPU path consumes it: P A P' = L L'.
struct HostFactor {
    int n = 0;
    std::vector<int> perm;      // perm[i] = original index of permuted row i
    // CSC of L: sorted rows, diagonal FIRST in each column (== CSR of L', the "upper" operand)
    std::vector<int> colptr, rowidx;
    std::vector<double> vals;
    int64_t nnz() const { return colptr.empty() ? 0 : (int64_t) colptr[n]; }
};

// Opaque CHOLMOD state (common + factor) kept alive for CPU-mode solvers.
struct CholmodState;
void cholmod_state_free(CholmodState *);

// Factorises A (lower part used, stype=-1).  Throws std::invalid_argument("Matrix is not positive
// definite!") like the reference (src/cholesky_solver.cpp:241-242).
// If keep_state != nullptr the CHOLMOD factor is kept (CPU mode) and *out_factor stays empty;
// otherwise L is extracted into *out_factor and all CHOLMOD memory is released (GPU mode).
void factorize_spd(const CscMatrix &A, HostFactor *out_factor, CholmodState **keep_state);

// CPU solve through cholmod_solve(CHOLMOD_A): row-major (n, k) Float in/out.
// (reference: src/cholesky_solver.cpp:434-463)
template <typename Float>
void cholmod_state_solve(CholmodState *st, int n_rhs, const Float *b, Float *x);

// Transpose a lower-triangular CSC factor into CSR (rows sorted by column, diagonal LAST).
void csc_to_csr(int n, const std::vector<int> &colptr, const std::vector<int> &rowidx, const std::vector<double> &vals,
                std::vector<int> &rowptr, std::vector<int> &colidx, std::vector<double> &rvals);

}  // namespace chb
