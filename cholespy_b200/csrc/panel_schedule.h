// Host-side construction of the device image the sm_100a sweep kernels stream: the factor L re-laid
// out as a 2-D block structure over PANELS (<= 16 consecutive columns of one supernode), grouped into
// warp-sized TASKS with a level-ordered ticket list.
//
// This replaces, for the solve kernels, the plain CSR/CSC upload + per-row level list of the
// reference (reference: src/cholesky_solver.cpp:245-288, 299-376; kernels/kernels.cu:64-202).  The
// reference's level arrays are still computed (GpuSolver::analyze) and exported bit-exactly; the
// panel/task schedule below is an internal product derived from the same dependency structure at
// supernode granularity.
//
// Terminology
//   panel   <= 16 consecutive columns c0..c0+w-1 of L treated as one dense block column: a dense lower
//           triangle on the diagonal and one below-panel row set R (the union over its columns).
//           Columns of a fundamental supernode share R exactly; small neighbouring supernodes are
//           amalgamated ("relaxed"), their absent entries stored as explicit zeros.
//   block   the dense rectangle L[R ∩ rows(panel t), cols(panel p)]: "block (t, p)".
//   tile    the w x n_rhs slab of unknowns of one panel, held in registers by one warp.
//   task    what one worker (a CTA of 1-4 warps, each owning 32 RHS columns) executes for one ticket:
//           a LOCAL task (a complete subtree of small panels, solved entirely in shared memory), or
//           one panel, or one chunk of the blocks of a large panel (partial sums parked in a
//           scratch slot).
//   local   a task whose unknowns (<= local_rows rows, one contiguous index range: the factor is
//           postordered, so a complete subtree is contiguous) are loaded into shared memory once,
//           together with the <= local_ext already-solved external rows the backward sweep needs
//           (all inside the subtree root's below-panel row list), solved there record by record
//           with no global-memory traffic, and written back once.  The forward sweep of a
//           complete subtree has no external dependency at all.
//   slab    the byte stream of one task: 16-byte-aligned records (below), streamed through shared
//           memory in pages by 1-D TMA bulk copies; no record straddles a page boundary.
//
// Inverse super-blocks (fast mode).  The panels of one wide supernode form a dependency CHAIN: panel t
// needs panel t - 1, one global hop per 16 columns, and the elimination tree's top separators make
// that chain hundreds of hops long.  For every run J of <= inv_panels consecutive panels of one
// supernode the dense triangle T = L[J, J] is therefore inverted on the host (double precision) and
// each panel t of J gets two tasks: A(t) accumulates r[t] = rhs[t] - (blocks from outside J) into the
// residual vector r, B(t) computes x[t] = sum_{p in J} Tinv[t, p] r[p] from the residuals of J.  All
// A tasks of J run in parallel, then all B tasks: two hops per run instead of one per panel.
//
// Forward sweep (L y = P b): task of panel t PULLS from the panels p < t it has blocks with:
//   tile[t] = b[perm[rows(t)]] - sum_p L[rows(t), cols(p)] y[cols(p)], then the in-panel triangle.
// Backward sweep (L' x = y): task of panel p PULLS from the panels t > p owning its rows R:
//   tile[p] = y[cols(p)] - sum_t L[R ∩ rows(t), cols(p)]' x[R ∩ rows(t)], then the triangle.
// In both sweeps every unknown accumulates its products in exactly the reference's order
// (ascending columns forward, descending rows backward, fused multiply-subtract, IEEE division)
// when ScheduleOptions::exact_order is set.  The default (fast) schedule instead splits large panels
// into PARALLEL partial sums, consumes blocks in order of readiness and multiplies by reciprocal
// pivots: deterministic, same operation count, different rounding (within the stated tolerance).
#pragma once
#include <cstdint>
#include <vector>

#include "factor_host.h"

namespace chb {

constexpr int kPanelWidth = 16;

// Record types of the slab stream.  Header = 4 x u32: {type | len16 << 8, a, b, c}; len16 = record
// length in 16-byte units including the header.
enum RecordType : uint32_t {
    REC_NOP = 0,         // padding up to the next page
    REC_BEGIN_RHS = 1,   // a = row0, b = w.  tile = rhs rows (forward: payload perm[16], gather from b; backward: tmp rows)
    REC_BEGIN_ZERO = 2,  // a = row0, b = w.  tile = 0
    REC_BEGIN_SLOT = 3,  // a = row0, b = w | slot << 8, c = wait + 1.  tile = scratch slot (chained partial sums)
    REC_ADD_SLOT = 4,    // b = slot << 8, c = wait + 1.  tile += scratch slot
    REC_BLOCK_D4 = 5,    // a = src_row0, b = smask | tmask << 16, c = wait + 1.  "grouped source-major" values: per source
                         // (sweep order) 4 values for every 4-target group that holds a target, zero-filled inside
    REC_BLOCK_D1 = 6,    // (retired layouts, never emitted)
    REC_BLOCK_TM = 7,
    REC_END_FINAL = 8,   // a = row0, b = w.  payload: triangle in sweep order (+ perm[16] backward); solve, store, publish
    REC_END_SLOT = 9,    // b = slot << 8.  park the tile in a scratch slot
    REC_END_DENSE = 10,  // a = row0, b = w.  like END_FINAL with the triangle as a dense wp x wp matrix (wp = 1 or
                         // round4(w)) holding RECIPROCAL pivots: multiply instead of divide (fast mode only)
    REC_LOCAL_LOAD = 11, // a = row_lo, b = n_int | n_ext << 16, c = n_wait.  payload (int32 each, every list padded
                         // to 4 entries): forward perm[n_int]; backward ext_row[n_ext], wait_code[n_wait].
                         // local[0..n_int) = rhs rows (forward: b[perm[i]]; backward: tmp[row_lo + i]),
                         // local[n_int + e] = tmp[ext_row[e]] after every wait code has published
    REC_LOCAL_STORE = 12,// a = row_lo, b = n_int.  payload backward: perm[n_int].  tmp[row_lo + i] = local[i]
                         // (backward: also x[perm[i]] = local[i]); the task's last record
    REC_LOCAL_ROWS = 14, // a = n_rows, b = 1: pivots are reciprocals (multiply) / 0: diagonals (divide).  The rows of a local task
                         // as sparse rows, in sweep order; per row: 16-byte header {u32 local_row | n4 << 8, pad, pivot at
                         // byte 8}, 4 n4 local source indices (u8, padded to 16 bytes), 4 n4 values:
                         // local[row] = (local[row] - sum values[i] * local[idx[i]]) * or / pivot
    REC_ADD_ROWS = 15,   // a = first row of the contribution buffer, b = nt | tmask << 16, c = wait + 1: tile[t] += contribution
                         // row, for the nt targets of tmask in ascending order.  Forward sweep, fast mode: what a local task
                         // contributes to this panel, formed by the producer (its LOCAL_ROWS rows with bit 16 of the meta
                         // word set: accumulated from zero into local row n_int + e; LOCAL_STORE, c = first contribution
                         // row, b = n_int | n_ext << 16, writes them out)
    REC_END_STORE = 13,  // a = row0, b = w, c = destination: 0 = the work vector (payload backward: perm[16]; publishes the
                         // panel), 1 = the residual vector r of an inverse super-block.  No triangle.
    REC_SRC_R_FLAG = 0x80, // on a BLOCK record: the source rows are rows of the residual vector r, not of the work vector
    // (retired: local tasks used to carry panel records marked with this flag; they are sparse rows now, REC_LOCAL_ROWS)
    REC_LOCAL_FLAG = 0x40,
};

struct ScheduleOptions {
    int local_rows = 0;       // a complete subtree becomes one LOCAL task up to this many rows (0: 24)
    int local_ext = 0;        // ... this many external rows below its root panel (backward sweep; 0: 24 f32 / 31 f64)
    int local_cost = 65536;   // ... and this many stored values
    int split_cost = 4096;    // a panel with more values than this is split into chunk tasks
    int chunk_cost = 2048;    // values per chunk task
    double relax_alpha = 1.5; // relaxed panels: stored values <= alpha * non-zeros + beta (1.0 / 0: no fill)
    double relax_beta = 32.0;
    bool contributions = true;// forward sweep, fast mode: local tasks hand their parents ready-made partial sums (REC_ADD_ROWS)
    int inv_panels = 256;     // consecutive panels of one supernode whose diagonal block is applied as an explicit
                              // inverse (fast mode only; < 2: off).  See "inverse super-blocks" below.
    int inv_chunk = 8;        // inverse-apply tasks of >= 2 * inv_chunk blocks are split into parallel chunk tasks of inv_chunk
                              // blocks + a final task (0: never)
    bool exact_order = false; // true: chunks are CHAINED (reference summation order kept bit for bit)
                              // false: chunks run in parallel and are summed in a fixed order
};

struct TaskDesc {
    uint64_t offset;  // byte offset of the task's slab in the image
    uint32_t bytes;   // slab length (multiple of 16)
    uint32_t level;   // dependency level (informational on the device)
};

// The panel partition and block structure shared by both sweeps.
struct PanelStructure {
    int n = 0;
    int n_panels = 0;
    std::vector<int> c0;             // n_panels + 1: first column of each panel (c0[n_panels] = n)
    std::vector<int> panel_of_row;   // n
    std::vector<int> rows;           // concatenated below-panel row lists (sorted union over the panel's columns)
    std::vector<int64_t> rlo;        // n_panels: offset of the panel's list in `rows`
    std::vector<int> m;              // n_panels: length of that list
    std::vector<int> parent;         // n_panels: panel owning the first below-panel row (-1: root)
    // blocks enumerated per SOURCE panel p (columns), ascending target panel
    std::vector<int64_t> bptr;       // n_panels + 1
    std::vector<int> b_src, b_tgt;   // panel ids
    std::vector<int> b_q0;           // offset of the block's first row inside the source panel's row list
    std::vector<uint16_t> b_mask;    // rows of the target panel present in the block
    // the same blocks listed per TARGET panel (stable: ascending source panel)
    std::vector<int64_t> fptr;       // n_panels + 1
    std::vector<int64_t> flist;      // block ids
    int n_supernodes = 0;
    int width(int p) const { return c0[p + 1] - c0[p]; }
};

struct SweepImage {
    bool upper = false;
    int elem = 4;
    int n = 0;
    int page_bytes = 4096;
    std::vector<uint8_t> slab;
    std::vector<TaskDesc> tasks;
    std::vector<int32_t> order;      // ticket -> task id, sorted by level (dependencies first)
    int n_slots = 0;                 // scratch tiles (kPanelWidth x n_rhs elements each)
    int64_t n_contrib_rows = 0;      // rows of the contribution buffer (forward sweep)
    int max_local_rows = 0;          // largest n_int + n_ext over the local tasks (sizes the kernels' shared memory)
    int n_levels = 0;
    int64_t n_values = 0;            // factor values referenced (== nnz(L))
    int64_t n_stored_values = 0;     // including zero fill
    int64_t n_blocks = 0;
};

void build_panel_structure(const HostFactor &L, const ScheduleOptions &opt, PanelStructure *out);
void build_sweep_image(const HostFactor &L, const PanelStructure &S, int elem_size, bool upper,
                       const ScheduleOptions &opt, SweepImage *out);

inline int page_bytes_for(int) { return 4096; }
// Defaults, overridable for tuning / parity runs: CHB_EXACT_ORDER=1, CHB_LOCAL_ROWS, CHB_LOCAL_EXT,
// CHB_LOCAL_COST, CHB_SPLIT_COST, CHB_CHUNK_COST, CHB_RELAX_ALPHA, CHB_RELAX_BETA.
ScheduleOptions schedule_options_from_env();

}  // namespace chb
