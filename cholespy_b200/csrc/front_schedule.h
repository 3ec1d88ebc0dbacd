// Host-side construction of the "front" device image: the factor L re-laid out for the sm_100a front
// sweep kernels (front_sweep.cuh) as a multifrontal-style task graph over SUPERNODES.
//
// What it replaces: the plain CSR/CSC upload + per-row level list of the reference
// (reference: src/cholesky_solver.cpp:245-288, 299-376) as the operand of solve_lower / solve_upper
// (kernels/kernels.cu:64-202).  The reference's level arrays are still computed and exported bit-exactly
// (GpuSolver::analyze); this schedule is internal.  It is the ALTERNATIVE image (CHB_SWEEP=front): the panel/block
// image (panel_schedule.h) is the default, being faster on every workload measured so far.
//
// Why a second layout (profiles/r01z_ncu_sweep_cfg4_summary.txt): the panel/block kernels spent 79 % of
// their issue slots on record interpretation (16-row blocks of a few sources each, 1.6-1.75x zero fill)
// and re-fetched every solved row once per (source panel, target panel) pair.  Here
//   * columns with identical structure (maximal supernodes, no fill at all) form dense block columns;
//   * the bottom of the elimination tree (complete subtrees of <= local_rows rows, ~80 % of the rows and
//     ~22 % of the factor on the mesh problems) is solved inside one CTA's shared memory as sparse rows
//     (LOCAL tasks, unchanged idea from round 1, but 2-4x larger);
//   * everything above is DENSE: a supernode J (columns C_J, below-rows R_J) is
//        forward   v = b[C_J] - (update rows addressed to C_J);  y = inv(L_JJ) v;  U_J = L[R_J, C_J] y
//        backward  v = y[C_J] - L[R_J, C_J]' x[R_J];             x = inv(L_JJ)' v
//     with the diagonal blocks (<= panel columns) inverted on the host in double precision, so that
//     every dense step is a register-tiled  (16 rows x n_rhs) += A(16 x s) X(s x n_rhs)  product whose
//     factor values stream through shared memory once and whose X rows are staged once per CTA;
//   * a supernode's products for its ancestors are written as UPDATE ROWS (one row of n_rhs values per
//     below-row, `W` buffer) that the owner of the row subtracts when it assembles its right-hand side:
//     every update row is written once and read once (multifrontal traffic, sum_J |R_J| rows), instead
//     of one gather per (source, target) block pair;
//   * wide supernodes are cut into panels (one task per panel, chained through ready flags), their
//     below-products into row groups x column ranges that run in parallel (fixed summation order:
//     deterministic);
//   * tickets are ordered subtree block by subtree block (level order inside a block), so update rows
//     are consumed while still L2-resident.
//
// Record stream ("slab") of a task: 16-byte header {type | len16 << 8 (16 bits), a, b, c} + payload, no record
// straddles a page (page_bytes).  STAGED records (FR_LOADX) are executed by the CTA's loader warp: wait
// for producer flags, copy rows into the shared-memory row ring with TMA; all other records by the compute
// warps.  Staged records are numbered 0, 1, ... inside a task (`sj`) and released in that order.
#pragma once
#include <cstdint>
#include <vector>

#include "factor_host.h"
#include "panel_schedule.h"  // TaskDesc

namespace chb {

constexpr int kFrontMaxWarps = 8;     // most compute warps a CTA shape may have (a dense step handles one tile per warp)
constexpr int kFrontChunk = 16;       // source rows per MMA record
constexpr int kFrontLoad = 32;        // source rows per staged chunk (two MMA records)
constexpr int kFrontPageExtra = 256;  // page = 8 tiles x 16 sources x tile_rows x elem + this

inline int front_tile_rows(int elem) { return elem == 8 ? 8 : 16; }
inline int front_page_bytes(int elem, int warps) { return warps * kFrontChunk * front_tile_rows(elem) * elem + kFrontPageExtra; }

enum FrontRecord : uint32_t {
    FR_NOP = 0,
    // a = n_contig | n_gather << 10 | n_extra << 20;  b = first contiguous row;
    // c = contig buffer (0 none, 1 Y, 2 W) | gather kind << 4 (0 none, 1 rows of Y, 2 rows of the user's b through
    // perm) | n_wait << 8.  payload: wait[n_wait] task ids (padded to 4 ints), gather[n_gather] (padded to 4).
    // Ring rows of the record: [contiguous | gathered | extra (allocated, not loaded)].
    FR_LOADX = 1,
    // a = tile0 | n_tiles << 8 | n_src << 16 | zero_init << 24;  b = sj | first source row inside it << 16.
    // payload: for each tile, [n_src][tile_rows] factor values.  acc[tile] (+)= A X, warp tile0 + t owns tile t.
    FR_MMA = 2,
    // a = tile0 | n_tiles << 8 | mode << 16;  b = n_rows (valid rows over all tiles) | sj << 16;  c = first global row.
    // payload: [0] = first ring row inside sj, then perm[n_rows] when FE_DST.
    FR_EPI = 3,
    // a = sj | first ring row << 16;  b = n_entries.  payload: seg[n_warps + 1] (entry range of each warp, padded to 12 ints),
    // entries {ring row (relative), W row}: ring[row] -= W[w_row].  Warp w owns ring rows == w (mod n_warps).
    FR_ASM = 4,
    FR_BAR = 5,   // barrier of the compute warps
    FR_REL = 6,   // a = sj: the compute warps are done with that staged record
    // b = sj.  Sparse rows of ONE dependency level of a local task, dealt to the compute warps.  payload: seg[n_warps + 1]
    // (byte range of each warp's rows, relative to the payload; padded to 12 ints), then per row: {u32 local row |
    // n4 << 8 | contribution << 16, pivot at byte elem}, 4 n4 local source indices (u8) from byte 2 elem, padded to 16
    // bytes, 4 n4 values:  local[row] = ((contribution ? 0 : local[row]) - sum values[i] local[idx[i]]) * pivot
    FR_LROWS = 7,
    // a = n_int | n_ext << 16, b = sj, c = first Y row.  payload: [0] = first W row (forward), [4..] perm[n_int] (backward)
    FR_LSTORE = 8,
    // First record of every page (16 bytes): bits 24-31 of the header word = number of FR_LOADX records in the page
    // (7: more than six, walk the page), then six u16 offsets of those records, in 16-byte units from the page start.
    FR_DIR = 9,
    // a = T (<= 4 n_warps), b = sj.  The upper rows of a local task, coupled among themselves by a small triangular
    // system: x = inv(T) v as a dense product.  payload: T local row indices (u8, padded to 16 bytes); per warp w,
    // T x 4 coefficients [q][i] = inv[w + i n_warps][q].  Every warp forms its (<= 4) rows from the CURRENT values of
    // all T rows, a barrier, then stores them.
    FR_LTOP = 10,
};
enum FrontEpiMode : uint32_t {
    FE_RING_SUB = 1,  // ring[row] -= acc
    FE_RING_SET = 2,  // ring[row] = acc
    FE_Y = 4,         // Y[c + r] = acc
    FE_W = 8,         // W[c + r] = acc
    FE_DST = 16,      // x[perm[r]] = acc   (backward)
};

struct FrontOptions {
    int warps = 7;           // compute warps of the CTA shape the image is built for (7: 2 CTAs per SM; 4: 3 per SM)
    int local_rows = 0;      // rows of a local task (0: default for the element size)
    int local_ext = 0;       // external rows of a local task
    int panel = 0;           // columns per dense panel (0: one tile per compute warp)
    int fuse_steps = 48;     // a single-panel supernode keeps its below-product in the same task up to this many MMA records
    int split_chunks = 16;   // staged source chunks (of 32 rows) per partial-product task
    int own_chunks = 8;      // trailing staged source chunks a panel task accumulates itself
    int order_blocks = 0;    // subtree blocks of the ticket order (0: plain level order; measured: blocks leave CTAs polling)
};
FrontOptions front_options_from_env(int elem);

struct FrontImage {
    bool upper = false;
    int elem = 4;
    int n = 0;
    int tile_rows = 16;
    int n_warps = 7;
    int page_bytes = 0;
    std::vector<uint8_t> slab;
    std::vector<TaskDesc> tasks;     // by task id; `level` = dependency level
    std::vector<int32_t> order;      // ticket -> task id (a topological order)
    int64_t n_w_rows = 0;            // rows of the update-row buffer W
    int max_stage_rows = 0;          // largest staged record (sizes the row ring)
    int n_levels = 0;
    int64_t n_values = 0;            // factor entries referenced (== nnz(L); diagonal blocks count through their inverse)
    int64_t n_stored_values = 0;     // values in the slab
    int64_t n_local_tasks = 0, n_dense_tasks = 0, n_mma = 0, n_asm_entries = 0;
    int64_t n_local_levels = 0, n_local_rows = 0;  // dependency levels / sparse rows over all local tasks
    int64_t local_phase_groups[4] = {0, 0, 0, 0};   // critical-path entry groups (of 4) per phase kind: subtrees, upper A, upper B, contributions
    int64_t local_phase_rows[4] = {0, 0, 0, 0};     // ... and rows on that critical path
};

// Supernode partition + local tasks shared by both sweeps.
struct FrontStructure {
    int n = 0;
    int ns = 0;
    std::vector<int> c0;          // ns + 1: first column of each supernode
    std::vector<int> sn_of;       // n: supernode of a column
    std::vector<int> parent;      // ns: parent supernode (-1: root)
    std::vector<int> local_of;    // ns: local task index, or -1 (dense)
    struct Local {
        int sn_lo, sn_hi;         // supernodes [sn_lo, sn_hi]
        int c_lo, c_hi;           // columns [c_lo, c_hi)
        std::vector<int> ext;     // external rows (sorted union of the roots' below-rows)
        std::vector<int> roots;   // root supernodes of the task's subtrees
    };
    std::vector<Local> locals;
    std::vector<int> block_of;    // ns: subtree block of the ticket order (the last block = everything above)
    int n_blocks = 1;
    int width(int s) const { return c0[s + 1] - c0[s]; }
};

void build_front_structure(const HostFactor &L, int elem, const FrontOptions &opt, FrontStructure *out);
void build_front_image(const HostFactor &L, const FrontStructure &S, int elem, bool upper, const FrontOptions &opt,
                       FrontImage *out);

}  // namespace chb
