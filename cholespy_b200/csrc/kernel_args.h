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

// One doubling round of the elimination-tree analysis (kernels.cu: chb_etree_round).
struct EtreeArgs {
    int n;
    int step;             // 2^round: distance covered by `anc_in`
    const int *anc_in;    // 2^round-th ancestor of every node (-1: none)
    int *anc_out;         // 2^(round+1)-th ancestor
    const int *dist_in;   // number of tree edges from the node to anc_in (or to its root when anc_in == -1)
    int *dist_out;
    const int *best_in;   // longest downward path of < 2^round edges starting at the node
    int *best_out;        // ... of < 2^(round+1) edges; must be zeroed before the round
    int *active;          // incremented by every node that still has an ancestor pointer after the round
};

// One panel sweep (kernels.cu: chb_sweep_*).  See panel_schedule.h for the slab/record format.
// Device task table, in TICKET order (dependencies first): entry t is what the worker drawing ticket t runs.
struct SweepTask {
    unsigned long long offset;  // byte offset of the task's slab
    unsigned int bytes;
    unsigned int level;         // on the device: the task's id (index of its ready flag; wait codes are id + 1)
};

template <typename Float>
struct SweepArgs {
    int n_tasks;
    int k;
    int epoch;
    int local_rows;       // largest local task (informational; local tasks live in the row ring)
    int ring_rows;        // rows of the shared-memory row ring (>= largest local task + 32)
    int src_tma;          // rows of the user's b can be fetched by TMA (16-byte aligned, 16-byte multiples)
    int src_ld, dst_ld;   // elements between consecutive rows of b / x (== k unless a column block of a wider array is solved)
    const unsigned char *slab;
    const SweepTask *tasks;     // ticket-ordered
    int *done;            // per-task epoch flags
    int *ticket;          // this sweep's ticket counter
    int *ticket_next;     // the other sweep's counter: cleared here
    Float *tmp;           // (n, k) work vector in permuted order
    const Float *src;     // forward: user's b
    Float *dst;           // backward: user's x
    Float *scratch;       // n_slots tiles of 16 x round4(k)
    Float *rbuf;          // (n, round4(k)) residual vector of the inverse super-blocks
    Float *cbuf;          // (n_contrib_rows, round4(k)) forward contributions of the local tasks
    unsigned long long *trace;  // optional (debug): 8 words per task {start ns, slab ready ns, end ns, cycles spent waiting, last wait ns, blocks done ns, triangle done ns, -}
};

// The local tasks of one panel sweep as a plain launch (local_sweep.cuh: chb_local_*).
template <typename Float>
struct LocalArgs {
    int n_tasks;          // local tasks only
    int k;
    int epoch;
    int slab_cap;         // bytes of one shared-memory record buffer (>= the largest local task, a multiple of 16)
    int src_vec, dst_vec; // a lane's columns of a row of b / x can be moved as one vector (alignment of base and stride)
    int src_ld, dst_ld;   // elements between consecutive rows of b / x
    const unsigned char *slab;
    const SweepTask *tasks;     // the local tasks, ticket order; `level` = task id = index of its ready flag
    int *done;            // per-task epoch flags (forward: published for the persistent kernel that follows)
    Float *tmp;           // (n, round4(k)) work vector in permuted order
    const Float *src;     // forward: user's b
    Float *dst;           // backward: user's x
    Float *cbuf;          // forward contributions of the local tasks
};

// One wide dependency level of a panel sweep as a plain launch (wide_sweep.cuh: chb_wide_*).
template <typename Float>
struct WideArgs {
    int n_tasks;          // the level's tasks
    int k;
    int epoch;
    int src_vec, dst_vec; // a lane's columns of a row of b / x can be moved as one vector (alignment of base and stride)
    int src_ld, dst_ld;   // elements between consecutive rows of b / x
    const unsigned char *slab;
    const SweepTask *tasks;     // the level's tasks; `level` = task id = index of its ready flag
    int *done;            // per-task epoch flags (published for the launches that follow)
    Float *tmp;           // (n, round4(k)) work vector in permuted order
    const Float *src;     // forward: user's b
    Float *dst;           // backward: user's x
    Float *scratch;       // partial tiles of chunk tasks
    Float *rbuf;          // residual vector of the inverse super-blocks
    Float *cbuf;          // forward contributions of the local tasks
};

// One front sweep (front_sweep.cuh: chb_front_*).  See front_schedule.h for the record format.
template <typename Float>
struct FrontArgs {
    int n_tasks;
    int k;
    int epoch;
    int ring_rows;        // rows of the shared-memory row ring (>= the largest staged record)
    int page_bytes;       // slab page size (front_schedule.h: front_page_bytes)
    int src_tma;          // rows of the user's b can be fetched by TMA (16-byte aligned, 16-byte multiples)
    int dst_vec;          // rows of the user's x can be written with vector stores
    int src_ld, dst_ld;   // elements between consecutive rows of b / x (== k unless a column block of a wider array is solved)
    int debug;            // timing experiments only (CHB_FRONT_DEBUG): bits skip LROWS / LSTORE / BAR / MMA / ASM / EPI bodies (results invalid)
    const unsigned char *slab;
    const SweepTask *tasks;     // ticket-ordered; `level` = task id = index of its ready flag
    int *done;            // per-task epoch flags
    int *ticket;          // this sweep's ticket counter
    int *ticket_next;     // the other sweep's counter: cleared here
    int *error;           // raised by the watchdog when a dependency never arrives (0 = fine)
    Float *tmp;           // (n, round4(k)) work vector in permuted order: y after the forward sweep, x after the backward
    Float *wbuf;          // (n_w_rows, round4(k)) update rows / partial products
    const Float *src;     // forward: user's b;  backward: null
    Float *dst;           // backward: user's x;  forward: null
    unsigned long long *trace;  // optional (debug): 8 words per task {start ns, CTA, end ns, first staged rows ready ns,
                                // ns warp 0 spent waiting for staged rows, loader: task start ns, first staged record issued ns,
                                // task done ns}
};

}  // namespace chb
