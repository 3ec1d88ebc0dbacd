// Kernel parameter blocks shared by kernels.cu (device) and gpu_solver.cpp (host, passed by value
// through cuLaunchKernel).  Plain data only.
#pragma once

namespace chb {

// Level-set analysis of one triangular CSR operand (kernels.cu: chb_analysis).
struct AnalysisArgs {
    int n;
    int descending;      // 0: CSR of L (deps = all but the last entry); 1: CSR of L' (all but the first)
    const int *rowptr;
    const int *colidx;
    int *level;          // pre-set to -1; the level word doubles as the ready flag
    int *max_level;
    int *ticket;
};

// One triangular sweep (kernels.cu: chb_solve_*).
template <typename Float>
struct SolveArgs {
    int n;
    int k;
    int upper;
    int epoch;
    const int *order;    // row ids in level order
    const int *rowptr;
    const int *colidx;
    const Float *vals;
    int *ready;          // per-row epoch flags
    int *ticket;         // this sweep's ticket counter
    int *ticket_next;    // the other counter: cleared here for the next sweep
    Float *tmp;          // (n, k) work vector in permuted order
    const Float *src;    // lower: user's b
    Float *dst;          // upper: user's x
    const int *perm;
};

}  // namespace chb
