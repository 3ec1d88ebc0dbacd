// Panel / block / task schedule of the factor for the sm_100a sweep kernels.  See panel_schedule.h.
#include "panel_schedule.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace chb {

namespace {

inline size_t round16(size_t b) { return (b + 15) & ~size_t(15); }
inline int popc16(uint32_t m) { return __builtin_popcount(m); }
inline int span16(uint32_t m) { return m ? 32 - __builtin_clz(m) : 0; }  // highest set bit + 1

// Appends records to the image, keeping every record inside one page of its task's slab.
class SlabWriter {
public:
    SlabWriter(std::vector<uint8_t> &buf, int page, int elem) : m_buf(buf), m_page((size_t) page), m_elem(elem) {}
    void begin_task() { m_task_start = m_buf.size(); }
    size_t size() const { return m_buf.size(); }

    // Reserves a zero-filled record; returns the byte offset of its payload.
    size_t record(uint32_t type, int32_t a, uint32_t b, uint32_t c, size_t payload_bytes) {
        const size_t len = 16 + round16(payload_bytes);
        if (len > m_page) throw std::logic_error("panel schedule: record larger than a page");
        const size_t in_page = (m_buf.size() - m_task_start) % m_page;
        if (in_page + len > m_page) header(REC_NOP, 0, 0, 0, m_page - in_page);
        const size_t off = m_buf.size();
        header(type, a, b, c, len);
        return off + 16;
    }
    void put_value(size_t payload_off, size_t index, double v) {
        if (m_elem == 4) {
            const float f = (float) v;  // host-side cast, as reference src/cholesky_solver.cpp:267-270
            std::memcpy(&m_buf[payload_off + 4 * index], &f, 4);
        } else {
            std::memcpy(&m_buf[payload_off + 8 * index], &v, 8);
        }
    }
    void put_int(size_t payload_off, size_t index, int32_t v) { std::memcpy(&m_buf[payload_off + 4 * index], &v, 4); }
    void put_bytes(size_t payload_off, const void *src, size_t n) { std::memcpy(&m_buf[payload_off], src, n); }

private:
    void header(uint32_t type, int32_t a, uint32_t b, uint32_t c, size_t len) {
        const size_t off = m_buf.size();
        m_buf.resize(off + len);  // zero-filled
        const uint32_t h[4] = {type | (uint32_t) (len / 16) << 8, (uint32_t) a, b, c};
        std::memcpy(&m_buf[off], h, 16);
    }
    std::vector<uint8_t> &m_buf;
    size_t m_page;
    int m_elem;
    size_t m_task_start = 0;
};

}  // namespace

ScheduleOptions schedule_options_from_env() {
    ScheduleOptions o;
    auto geti = [](const char *name, int dflt) {
        const char *v = std::getenv(name);
        return v && *v ? std::atoi(v) : dflt;
    };
    o.exact_order = geti("CHB_EXACT_ORDER", 0) != 0;
    o.local_rows = std::max(0, std::min(240, geti("CHB_LOCAL_ROWS", o.local_rows)));
    o.local_ext = std::max(0, std::min(31, geti("CHB_LOCAL_EXT", o.local_ext)));  // one loader lane per external row
    o.local_cost = std::max(1, geti("CHB_LOCAL_COST", o.local_cost));
    o.inv_panels = std::max(0, std::min(512, geti("CHB_INV_PANELS", o.inv_panels)));
    o.inv_chunk = std::max(0, geti("CHB_INV_CHUNK", o.inv_chunk));
    o.contributions = geti("CHB_CONTRIBUTIONS", o.contributions ? 1 : 0) != 0;
    o.split_cost = std::max(256, geti("CHB_SPLIT_COST", o.split_cost));
    o.chunk_cost = std::max(256, geti("CHB_CHUNK_COST", o.chunk_cost));
    if (const char *v = std::getenv("CHB_RELAX_ALPHA")) o.relax_alpha = std::max(1.0, std::atof(v));
    if (const char *v = std::getenv("CHB_RELAX_BETA")) o.relax_beta = std::max(0.0, std::atof(v));
    return o;
}

void build_panel_structure(const HostFactor &L, const ScheduleOptions &opt, PanelStructure *out) {
    PanelStructure &S = *out;
    const int n = L.n;
    const std::vector<int> &cp = L.colptr, &ri = L.rowidx;
    S = PanelStructure();
    S.n = n;
    S.panel_of_row.resize(n);
    S.c0.reserve(n / 4 + 2);

    // Greedy RELAXED panels: consecutive columns are packed into one panel (<= 16 wide) as long as the
    // dense layout it implies -- full triangle + (union of the columns' below-panel rows) x width,
    // absent entries stored as explicit zeros -- stays within alpha * (true non-zeros) + beta values.
    // Columns of one fundamental supernode have identical structure (no fill at all), so wide
    // supernodes simply split into 16-column panels; the many 1-3 column supernodes at the bottom
    // of the elimination tree are amalgamated, which is what keeps the number of panels, blocks
    // and dependency hops (the latency-bound part of the sweeps) small.  Explicit zeros are
    // bit-neutral under fused multiply-subtract, so the reference summation order is unaffected.
    //
    // A panel must not couple unrelated parts of the elimination tree, or it would chain sibling
    // subtrees through false dependencies: the columns of a panel whose etree parent lies outside it
    // ("tops") must all have the SAME parent.  Panels are therefore runs of complete sibling pieces,
    // optionally continued up their common parent chain.
    auto etree_parent = [&](int j) { return cp[j + 1] - cp[j] > 1 ? ri[cp[j] + 1] : -1; };
    std::vector<int> uni, merged;
    int a = 0;
    while (a < n) {
        int w = 1;
        int64_t true_vals = cp[a + 1] - cp[a];
        int tops_parent = etree_parent(a);
        uni.assign(ri.begin() + cp[a] + 1, ri.begin() + cp[a + 1]);  // rows > a of the panel so far (incl. in-panel)
        while (a + w < n && w < kPanelWidth) {
            const int jn = a + w;
            if (tops_parent != jn && tops_parent != etree_parent(jn)) break;  // would join unrelated subtrees
            merged.clear();
            std::set_union(uni.begin(), uni.end(), ri.begin() + cp[jn] + 1, ri.begin() + cp[jn + 1], std::back_inserter(merged));
            const int64_t tv = true_vals + (cp[jn + 1] - cp[jn]);
            const int wn = w + 1;
            const int64_t below = (int64_t) (merged.end() - std::upper_bound(merged.begin(), merged.end(), jn));
            const int64_t padded = (int64_t) wn * (wn + 1) / 2 + below * wn;
            if ((double) padded > opt.relax_alpha * (double) tv + opt.relax_beta) break;
            uni.swap(merged);
            true_vals = tv;
            w = wn;
            tops_parent = etree_parent(jn);  // either all old tops were children of jn, or they are its siblings
        }
        const int p = (int) S.c0.size();
        S.c0.push_back(a);
        for (int r = a; r < a + w; ++r) S.panel_of_row[r] = p;
        S.rlo.push_back((int64_t) S.rows.size());
        const auto first_below = std::lower_bound(uni.begin(), uni.end(), a + w);
        S.rows.insert(S.rows.end(), first_below, uni.end());
        S.m.push_back((int) (uni.end() - first_below));
        a += w;
    }
    S.n_panels = (int) S.c0.size();
    S.c0.push_back(n);
    const int np = S.n_panels;
    S.bptr.assign(np + 1, 0);
    S.parent.resize(np);
    for (int p = 0; p < np; ++p) S.parent[p] = S.m[p] > 0 ? S.panel_of_row[S.rows[S.rlo[p]]] : -1;
    const std::vector<int> &rr = S.rows;
    // blocks per source panel
    for (int p = 0; p < np; ++p) {
        const int64_t lo = S.rlo[p];
        const int m = S.m[p];
        int q = 0;
        while (q < m) {
            const int t = S.panel_of_row[rr[lo + q]];
            const int base = S.c0[t];
            uint32_t mask = 0;
            const int q0 = q;
            while (q < m && S.panel_of_row[rr[lo + q]] == t) {
                mask |= 1u << (rr[lo + q] - base);
                ++q;
            }
            S.b_src.push_back(p);
            S.b_tgt.push_back(t);
            S.b_q0.push_back(q0);
            S.b_mask.push_back((uint16_t) mask);
        }
        S.bptr[p + 1] = (int64_t) S.b_src.size();
    }
    // the same blocks bucketed by target panel (stable => ascending source panel inside a bucket)
    const int64_t nb = (int64_t) S.b_src.size();
    S.fptr.assign(np + 1, 0);
    for (int64_t b = 0; b < nb; ++b) S.fptr[S.b_tgt[b] + 1] += 1;
    for (int p = 0; p < np; ++p) S.fptr[p + 1] += S.fptr[p];
    S.flist.resize(nb);
    {
        std::vector<int64_t> cursor(S.fptr.begin(), S.fptr.end() - 1);
        for (int64_t b = 0; b < nb; ++b) S.flist[cursor[S.b_tgt[b]]++] = b;
    }
}

namespace {

struct Builder {
    const HostFactor &L;
    const PanelStructure &S;
    const ScheduleOptions &opt;
    SweepImage &img;
    SlabWriter w;
    const bool upper;
    std::vector<int> task_of_panel;
    int cur = -1;
    uint32_t cur_level = 0;
    // local task being emitted (see REC_LOCAL_LOAD): unknowns [loc_lo, loc_lo + loc_n_int), then the external rows
    int loc_lo = 0, loc_n_int = 0;
    std::vector<int> loc_ext;
    // inverse super-blocks: A tasks end by storing the residual instead of solving the triangle
    std::vector<int> local_root;  // root panel of the local subtree a panel belongs to, or -1
    // forward contributions of local tasks (see REC_ADD_ROWS): per local task its external target rows (sorted) and
    // the row of the contribution buffer its first contribution goes to
    struct Contribution {
        int64_t base = 0;
        std::vector<int> rows;
    };
    std::vector<Contribution> contrib_of_task;
    bool contributions() const { return !upper && !opt.exact_order && opt.contributions; }
    int64_t stat_blocks_from_local = 0, stat_rows_from_local = 0, stat_vals_from_local = 0, stat_blocks_nonlocal = 0, stat_rows_nonlocal = 0, stat_ext_rows = 0;
    double stat_inverse_s = 0, stat_inverse_panel_s = 0, stat_inverse_apply_s = 0, stat_local_s = 0, stat_fill_s = 0, stat_prep_s = 0;  // host seconds (CHB_SCHED_STATS)
    bool end_store_r = false;
    std::vector<int> atask_of_panel;

    Builder(const HostFactor &L_, const PanelStructure &S_, const ScheduleOptions &o, SweepImage &i)
        : L(L_), S(S_), opt(o), img(i), w(i.slab, i.page_bytes, i.elem), upper(i.upper), task_of_panel(S_.n_panels, -1), atask_of_panel(S_.n_panels, -1) {}

    // -- tasks ----------------------------------------------------------------------------------
    int open_task() {
        cur = (int) img.tasks.size();
        img.tasks.push_back(TaskDesc{(uint64_t) w.size(), 0u, 0u});
        w.begin_task();
        cur_level = 0;
        return cur;
    }
    void close_task() {
        if (cur < 0) return;
        const uint64_t bytes = (uint64_t) w.size() - img.tasks[cur].offset;
        if (bytes >= (1ull << 32)) throw std::logic_error("panel schedule: task slab too large");
        img.tasks[cur].bytes = (uint32_t) bytes;
        img.tasks[cur].level = cur_level;
        cur = -1;
    }
    uint32_t wait_code(int dep_task) {
        if (dep_task < 0) throw std::logic_error("panel schedule: dependency on an unscheduled panel");
        if (dep_task == cur) return 0;  // produced earlier by this same warp
        cur_level = std::max(cur_level, img.tasks[dep_task].level + 1);
        return (uint32_t) dep_task + 1;
    }

    // -- factor access --------------------------------------------------------------------------
    // Dense copy of L[R_p[q0 .. q0+nr), columns of p] (absent entries = 0); returns the number present.
    int gather_block(int p, int q0, int nr, double (&buf)[kPanelWidth][kPanelWidth]) const {
        const int c0 = S.c0[p], wd = S.width(p);
        const int *R = &S.rows[(size_t) S.rlo[p] + (size_t) q0];
        int present = 0;
        for (int j = 0; j < wd; ++j) {
            const int c = c0 + j;
            const int *cr = &L.rowidx[(size_t) L.colptr[c]];
            const int len = L.colptr[c + 1] - L.colptr[c];
            int at = (int) (std::lower_bound(cr, cr + len, R[0]) - cr);
            for (int q = 0; q < nr; ++q) {
                while (at < len && cr[at] < R[q]) ++at;
                if (at < len && cr[at] == R[q]) {
                    buf[q][j] = L.vals[(size_t) L.colptr[c] + (size_t) at];
                    present += 1;
                } else {
                    buf[q][j] = 0.0;
                }
            }
        }
        return present;
    }
    // L[c0(p) + i, c0(p) + j], i >= j (0 if absent; `present` counts the entries that exist)
    double tri(int p, int i, int j, int *present = nullptr) const {
        const int c = S.c0[p] + j, row = S.c0[p] + i;
        for (int e = L.colptr[c]; e < L.colptr[c + 1] && L.rowidx[e] <= row; ++e)
            if (L.rowidx[e] == row) {
                if (present) *present += 1;
                return L.vals[(size_t) e];
            }
        return 0.0;
    }

    int64_t block_cost(int64_t b) const { return (int64_t) popc16(S.b_mask[b]) * S.width(S.b_src[b]); }
    int64_t tri_cost(int p) const {
        const int wd = S.width(p);
        return (int64_t) wd * (wd + 1) / 2;
    }

    // Blocks of panel P in the order this sweep consumes them.
    void blocks_of(int P, std::vector<int64_t> &out) const {
        out.clear();
        if (upper) {
            for (int64_t b = S.bptr[P + 1] - 1; b >= S.bptr[P]; --b) out.push_back(b);  // farthest owner first
        } else {
            for (int64_t f = S.fptr[P]; f < S.fptr[P + 1]; ++f) out.push_back(S.flist[f]);  // ascending source
        }
    }

    // Unless the reference's summation order is requested, consume the blocks whose sources finish
    // earliest first (by dependency level), so that a task is left with as little work as possible
    // once its last dependency arrives.  Deterministic: a pure function of the factor's pattern.
    void order_by_readiness(std::vector<int64_t> &blocks) const {
        if (opt.exact_order) return;
        auto key = [&](int64_t b) -> int64_t {
            const int t = task_of_panel[upper ? S.b_tgt[b] : S.b_src[b]];
            return t == cur || t < 0 ? -1 : (int64_t) img.tasks[t].level;
        };
        std::stable_sort(blocks.begin(), blocks.end(), [&](int64_t x, int64_t y) { return key(x) < key(y); });
    }

    // -- records --------------------------------------------------------------------------------
    int local_index(int row) const {
        if (row >= loc_lo && row < loc_lo + loc_n_int) return row - loc_lo;
        const auto it = std::lower_bound(loc_ext.begin(), loc_ext.end(), row);
        if (it == loc_ext.end() || *it != row) throw std::logic_error("panel schedule: local task reads a row it did not load");
        return loc_n_int + (int) (it - loc_ext.begin());
    }

    void emit_begin_rhs(int P) {
        const int wd = S.width(P), row0 = S.c0[P];
        if (upper) {
            w.record(REC_BEGIN_RHS, row0, (uint32_t) wd, 0, 0);
        } else {
            const size_t off = w.record(REC_BEGIN_RHS, row0, (uint32_t) wd, 0, 4 * kPanelWidth);
            for (int i = 0; i < wd; ++i) w.put_int(off, (size_t) i, L.perm[row0 + i]);
        }
    }
    void emit_begin_zero(int P) { w.record(REC_BEGIN_ZERO, S.c0[P], (uint32_t) S.width(P), 0, 0); }
    void emit_begin_slot(int P, int slot, int dep_task) {
        w.record(REC_BEGIN_SLOT, S.c0[P], (uint32_t) S.width(P) | (uint32_t) slot << 8, wait_code(dep_task), 0);
    }
    void emit_add_slot(int slot, int dep_task) { w.record(REC_ADD_SLOT, 0, (uint32_t) slot << 8, wait_code(dep_task), 0); }
    void emit_end_slot(int slot) { w.record(REC_END_SLOT, 0, (uint32_t) slot << 8, 0, 0); }

    // A block record's header and masks are functions of the structure alone; its values (a sparse -> dense gather and
    // a re-layout) are the expensive part of building an image and independent of everything else: emit_block
    // reserves the record, fill_block writes the payload -- for all blocks at the end of run(), by all host threads.
    struct PendingBlock {
        size_t off;
        int64_t b;
    };
    std::vector<PendingBlock> pending_blocks;

    void emit_block(int P, int64_t b) {
        const int p = S.b_src[b], t = S.b_tgt[b];
        const uint32_t rows = S.b_mask[b];
        const int nr = popc16(rows);
        const size_t es = (size_t) img.elem;
        img.n_blocks += 1;
        // One layout for every block ("grouped source-major"): per source, in sweep order, the values of the
        // 4-target GROUPS that contain at least one target (zero-filled inside a group).  Explicit zeros
        // are bit-neutral under fused multiply-subtract.
        const int wd = S.width(p);
        const uint32_t tmask = upper ? (1u << wd) - 1 : rows;   // unknowns: columns of p (backward) / rows of t present (forward)
        const uint32_t smask = upper ? rows : (1u << wd) - 1;   // sources: rows of t present (backward) / columns of p (forward)
        const int ns = upper ? nr : wd;
        int ng = 0;
        for (int g = 0; g < 4; ++g) ng += (tmask >> (4 * g) & 0xf) ? 1 : 0;
        const int stride = 4 * ng;
        const uint32_t wait = wait_code(task_of_panel[upper ? t : p]);
        (void) P;
        if (!upper && !local_root.empty() && local_root[p] >= 0) stat_blocks_from_local += 1, stat_rows_from_local += ns, stat_vals_from_local += (int64_t) ns * popc16(tmask);
        if (!upper) stat_blocks_nonlocal += 1, stat_rows_nonlocal += ns;
        const size_t off = w.record(REC_BLOCK_D4, S.c0[upper ? t : p], smask | tmask << 16, wait, es * (size_t) ns * (size_t) stride);
        pending_blocks.push_back(PendingBlock{off, b});
        img.n_stored_values += (int64_t) ns * stride;
    }
    // The payload of one reserved block record; returns the number of factor entries it holds.  Touches only its own
    // bytes of the slab (which no longer grows when this runs).
    int fill_block(const PendingBlock &pb) {
        const int64_t b = pb.b;
        const int p = S.b_src[b], q0 = S.b_q0[b];
        const uint32_t rows = S.b_mask[b];
        const int nr = popc16(rows);
        double buf[kPanelWidth][kPanelWidth];  // [row of the block][column of p]
        const int present = gather_block(p, q0, nr, buf);
        const int wd = S.width(p);
        const uint32_t tmask = upper ? (1u << wd) - 1 : rows;
        const int ns = upper ? nr : wd;
        int group_rank[4], ng = 0;
        for (int g = 0; g < 4; ++g) group_rank[g] = (tmask >> (4 * g) & 0xf) ? ng++ : -1;
        const int stride = 4 * ng;
        // backward: sources visited in DESCENDING row order (the reference walks a row of L' from its end backwards)
        for (int r = 0; r < ns; ++r) {
            int q = 0;  // rank of the target among the block's rows (forward)
            for (int tt = 0; tt < kPanelWidth; ++tt) {
                if (!(tmask >> tt & 1)) continue;
                const double v = upper ? buf[nr - 1 - r][tt] : buf[q++][r];
                w.put_value(pb.off, (size_t) r * stride + 4 * group_rank[tt >> 2] + (tt & 3), v);
            }
        }
        return present;
    }
    void fill_pending_blocks() {
        const size_t n = pending_blocks.size();
        std::atomic<size_t> next{0};
        std::atomic<int64_t> present{0};
        constexpr size_t kBatch = 256;
        auto worker = [&]() {
            int64_t mine = 0;
            for (;;) {
                const size_t lo = next.fetch_add(kBatch);
                if (lo >= n) break;
                const size_t hi = std::min(n, lo + kBatch);
                for (size_t i = lo; i < hi; ++i) mine += fill_block(pending_blocks[i]);
            }
            present += mine;
        };
        const unsigned n_threads = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));  // both sweeps are built at once
        if (n_threads <= 1 || n < 4 * kBatch) {
            worker();
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < n_threads; ++t) pool.emplace_back(worker);
            for (std::thread &t : pool) t.join();
        }
        img.n_values += present.load();
        pending_blocks.clear();
        pending_blocks.shrink_to_fit();
    }

    // Fast mode: the triangle as a dense wp x wp matrix (wp = 1 or round4(w)) in pivot order with
    // RECIPROCAL pivots, so the kernel's substitution is branch-free with static shared-memory
    // offsets and has no division on its dependency chain.  Row s = pivot s: [s][s] = 1/L[s,s],
    // [s][t] = the multiplier of unknown t (forward: t > s, L[t,s]; backward: t < s, L[s,t]).
    void emit_end_dense(int P) {
        const int wd = S.width(P), row0 = S.c0[P];
        const size_t es = (size_t) img.elem;
        const int wp = wd == 1 ? 1 : (wd + 3) & ~3;
        const size_t head = upper ? 4 * kPanelWidth : 0;
        const size_t off = w.record(REC_END_DENSE, row0, (uint32_t) wd, 0, head + es * (size_t) (wp * wp));
        if (head)
            for (int i = 0; i < wd; ++i) w.put_int(off, (size_t) i, L.perm[row0 + i]);
        int present = 0;
        for (int s = 0; s < wp; ++s) {
            w.put_value(off + head, (size_t) s * wp + s, s < wd ? 1.0 / tri(P, s, s, &present) : 1.0);
            if (s >= wd) continue;
            if (!upper)
                for (int t = s + 1; t < wd; ++t) w.put_value(off + head, (size_t) s * wp + t, tri(P, t, s, &present));
            else
                for (int t = 0; t < s; ++t) w.put_value(off + head, (size_t) s * wp + t, tri(P, s, t, &present));
        }
        img.n_values += present;
        img.n_stored_values += (int64_t) wp * wp;
        task_of_panel[P] = cur;
    }

    // END_STORE: the tile goes to the residual vector (A task of an inverse super-block) or, as the
    // solved unknowns, to the work vector (B task)
    void emit_end_store(int P, bool to_residual) {
        const int wd = S.width(P), row0 = S.c0[P];
        const size_t head = upper && !to_residual ? 4 * kPanelWidth : 0;
        const size_t off = w.record(REC_END_STORE, row0, (uint32_t) wd, to_residual ? 1u : 0u, head);
        if (head)
            for (int i = 0; i < wd; ++i) w.put_int(off, (size_t) i, L.perm[row0 + i]);
        if (to_residual) atask_of_panel[P] = cur;
        else task_of_panel[P] = cur;
    }

    // A dense block of an inverse super-block: tile -= V * r[rows of panel p]; value(r, t) = V[target t][source
    // r-th in sweep order].
    template <typename ValueFn>
    void emit_inverse_block(int p, int n_targets, int dep_task, ValueFn value) {
        const int ns = S.width(p);
        const uint32_t smask = (1u << ns) - 1, tmask = (1u << n_targets) - 1;
        const int stride = (n_targets + 3) & ~3;
        const size_t off = w.record(REC_BLOCK_D4 | REC_SRC_R_FLAG, S.c0[p], smask | tmask << 16, wait_code(dep_task),
                                    (size_t) img.elem * (size_t) ns * (size_t) stride);
        for (int r = 0; r < ns; ++r)
            for (int t = 0; t < n_targets; ++t) w.put_value(off, (size_t) r * stride + t, value(r, t));
        img.n_blocks += 1;
        img.n_stored_values += (int64_t) ns * stride;
    }

    void emit_end_final(int P) {
        if (end_store_r) return emit_end_store(P, true);
        if (!opt.exact_order) return emit_end_dense(P);
        const int wd = S.width(P), row0 = S.c0[P];
        const size_t es = (size_t) img.elem;
        const size_t tri_bytes = round16(es * (size_t) (wd * (wd + 1) / 2));
        const size_t head = upper ? 4 * kPanelWidth : 0;  // backward: perm[16] first (x is scattered through it)
        const size_t off = w.record(REC_END_FINAL, row0, (uint32_t) wd, 0, head + tri_bytes);
        if (head)
            for (int i = 0; i < wd; ++i) w.put_int(off, (size_t) i, L.perm[row0 + i]);
        size_t idx = 0;
        int present = 0;
        if (!upper) {
            for (int s = 0; s < wd; ++s)  // pivot s, then column s of the triangle below it
                for (int t = s; t < wd; ++t) w.put_value(off + head, idx++, tri(P, t, s, &present));
        } else {
            for (int s = wd - 1; s >= 0; --s) {  // pivot s, then row s of the triangle left of it
                w.put_value(off + head, idx++, tri(P, s, s, &present));
                for (int t = 0; t < s; ++t) w.put_value(off + head, idx++, tri(P, s, t, &present));
            }
        }
        img.n_values += present;
        img.n_stored_values += (int64_t) idx;
        task_of_panel[P] = cur;
    }

    // One whole panel inside the current task.
    // Forward sweep, fast mode: what the local tasks below contribute to panel P arrives as ready-made partial sums
    // (one contribution row per target row and producer task, REC_ADD_ROWS) instead of one block per source panel:
    // the products are formed by the producer while its unknowns sit in shared memory.  Removes those blocks from
    // `blocks` and emits the additions.
    std::vector<int> take_contributions(std::vector<int64_t> &blocks) {
        std::vector<int> producers;
        if (!contributions()) return producers;
        size_t keep = 0;
        for (int64_t b : blocks) {
            const int src = S.b_src[b];
            if (local_root[src] >= 0) {
                const int t = task_of_panel[src];
                if (producers.empty() || producers.back() != t) producers.push_back(t);  // a task's panels are consecutive
            } else {
                blocks[keep++] = b;
            }
        }
        blocks.resize(keep);
        return producers;
    }
    void emit_add_rows(int P, const std::vector<int> &producers) {
        const int row_lo = S.c0[P], row_hi = S.c0[P + 1];
        for (int t : producers) {
            const Contribution &c = contrib_of_task[(size_t) t];
            const auto first = std::lower_bound(c.rows.begin(), c.rows.end(), row_lo);
            const auto last = std::lower_bound(c.rows.begin(), c.rows.end(), row_hi);
            const int nt = (int) (last - first);
            if (nt <= 0) throw std::logic_error("panel schedule: a producer without contribution rows");
            uint32_t tmask = 0;
            for (auto it = first; it != last; ++it) tmask |= 1u << (*it - row_lo);
            const int64_t crow = c.base + (int64_t) (first - c.rows.begin());
            if (crow >= (int64_t) 1 << 31) throw std::logic_error("panel schedule: contribution buffer too large");
            w.record(REC_ADD_ROWS, (int32_t) crow, (uint32_t) nt | tmask << 16, wait_code(t), 0);
        }
    }

    void emit_panel(int P, std::vector<int64_t> &blocks) {
        order_by_readiness(blocks);
        const std::vector<int> producers = take_contributions(blocks);
        emit_begin_rhs(P);
        emit_add_rows(P, producers);
        for (int64_t b : blocks) emit_block(P, b);
        emit_end_final(P);
    }

    // A large panel: chunk tasks + a final task.
    void emit_split_panel(int P, std::vector<int64_t> &blocks) {
        order_by_readiness(blocks);
        const std::vector<int> producers = take_contributions(blocks);
        std::vector<size_t> cut{0};  // chunk boundaries in `blocks`
        int64_t acc = 0;
        for (size_t i = 0; i < blocks.size(); ++i) {
            acc += block_cost(blocks[i]);
            if (acc >= opt.chunk_cost && i + 1 < blocks.size()) {
                cut.push_back(i + 1);
                acc = 0;
            }
        }
        cut.push_back(blocks.size());
        const int n_chunks = (int) cut.size() - 1;
        std::vector<int> slot(n_chunks, -1), task(n_chunks, -1);
        for (int c = 0; c + 1 < n_chunks; ++c) {
            task[c] = open_task();
            if (opt.exact_order) {
                if (c == 0) emit_begin_rhs(P);
                else emit_begin_slot(P, slot[c - 1], task[c - 1]);
            } else {
                emit_begin_zero(P);
            }
            for (size_t i = cut[c]; i < cut[c + 1]; ++i) emit_block(P, blocks[i]);
            slot[c] = img.n_slots++;
            emit_end_slot(slot[c]);
            close_task();
        }
        open_task();
        if (n_chunks == 1) {
            emit_begin_rhs(P);
            emit_add_rows(P, producers);
        } else if (opt.exact_order) {
            emit_begin_slot(P, slot[n_chunks - 2], task[n_chunks - 2]);
        } else {
            emit_begin_rhs(P);
            emit_add_rows(P, producers);
            for (int c = 0; c + 1 < n_chunks; ++c) emit_add_slot(slot[c], task[c]);
        }
        for (size_t i = cut[n_chunks - 1]; i < cut[n_chunks]; ++i) emit_block(P, blocks[i]);
        emit_end_final(P);
        close_task();
    }

    // The rows of a local task as plain sparse rows (REC_LOCAL_ROWS): at the bottom of the elimination tree rows hold
    // a handful of non-zeros, and a row-by-row substitution out of shared memory costs a fraction of the panel /
    // block machinery.  Forward: row i of L, ascending columns; backward: column j of L (row j of L'), descending
    // rows -- the reference's summation orders.  Entries come in groups of 4 (zero padding is bit-neutral).
    void emit_local_rows(int lo, int hi) {
        const int n_int = loc_n_int, c_lo = loc_lo;
        struct Entry {
            int idx;
            double v;
        };
        std::vector<std::vector<Entry>> rows((size_t) n_int);
        std::vector<double> diag((size_t) n_int);
        for (int j = 0; j < n_int; ++j) {
            const int c = c_lo + j;
            diag[j] = L.vals[(size_t) L.colptr[c]];
            if (!upper) {
                for (int e = L.colptr[c] + 1; e < L.colptr[c + 1] && L.rowidx[e] < c_lo + n_int; ++e)
                    rows[(size_t) (L.rowidx[e] - c_lo)].push_back(Entry{j, L.vals[(size_t) e]});  // ascending columns per row
            } else {
                for (int e = L.colptr[c + 1] - 1; e > L.colptr[c]; --e)
                    rows[j].push_back(Entry{local_index(L.rowidx[e]), L.vals[(size_t) e]});       // descending rows
            }
        }
        (void) lo, (void) hi;
        // forward contributions: one more sparse row per external target row r, local row n_int + e, accumulated from
        // zero: -(sum over the task's columns c of L[r, c] x[c]); the target panel adds it (REC_ADD_ROWS)
        const int n_out = contributions() ? (int) loc_ext.size() : 0;
        if (n_out > 0) {
            rows.resize((size_t) (n_int + n_out));
            diag.resize((size_t) (n_int + n_out), 1.0);
            for (int j = 0; j < n_int; ++j) {
                const int c = c_lo + j;
                for (int e = L.colptr[c + 1] - 1; e > L.colptr[c] && L.rowidx[e] >= c_lo + n_int; --e)
                    rows[(size_t) local_index(L.rowidx[e])].push_back(Entry{j, L.vals[(size_t) e]});
            }
        }
        const int n_rows_total = n_int + n_out;
        const size_t es = (size_t) img.elem;
        auto row_bytes = [&](int j) {
            const size_t n4 = (rows[j].size() + 3) / 4;
            return 16 + round16(4 * n4) + 4 * n4 * es;
        };
        std::vector<int> sequence((size_t) n_rows_total);  // rows in processing order
        for (int q = 0; q < n_rows_total; ++q) sequence[(size_t) q] = upper ? n_int - 1 - q : q;
        int j = 0;
        while (j < n_rows_total) {
            // as many rows as fit one page (a record never straddles a page)
            size_t bytes = 0;
            int cnt = 0;
            while (j + cnt < n_rows_total && 16 + bytes + row_bytes(sequence[(size_t) (j + cnt)]) <= (size_t) img.page_bytes) {
                bytes += row_bytes(sequence[(size_t) (j + cnt)]);
                cnt += 1;
            }
            if (cnt == 0) throw std::logic_error("panel schedule: a local row does not fit a page");
            const size_t off = w.record(REC_LOCAL_ROWS, cnt, opt.exact_order ? 0u : 1u, 0, bytes);
            size_t at = off;
            for (int q = 0; q < cnt; ++q) {
                const int r = sequence[(size_t) (j + q)];
                const bool contribution = r >= n_int;  // accumulated from zero, no pivot
                const std::vector<Entry> &ent = rows[(size_t) r];
                const size_t n4 = (ent.size() + 3) / 4;
                if (n4 > 255) throw std::logic_error("panel schedule: local row too long");
                w.put_int(at, 0, (int32_t) ((uint32_t) r | (uint32_t) n4 << 8 | (contribution ? 1u << 16 : 0u)));
                // the pivot at byte 8: reciprocal in fast mode (multiply), the diagonal itself in exact-order mode (IEEE divide)
                w.put_value(at + 8, 0, contribution ? 1.0 : opt.exact_order ? diag[(size_t) r] : 1.0 / diag[(size_t) r]);
                uint8_t idx[1024];
                for (size_t i = 0; i < 4 * n4; ++i) idx[i] = (uint8_t) (i < ent.size() ? ent[i].idx : contribution ? 0 : r);
                w.put_bytes(at + 16, idx, 4 * n4);
                const size_t voff = at + 16 + round16(4 * n4);
                for (size_t i = 0; i < ent.size(); ++i) w.put_value(voff, i, ent[i].v);
                at = voff + 4 * n4 * es;
                img.n_values += (int64_t) ent.size() + (contribution ? 0 : 1);
                img.n_stored_values += (int64_t) (4 * n4) + 1;
            }
            j += cnt;
        }
    }

    // One LOCAL task: the complete subtree of panels [lo, hi] (hi = its root), solved in shared memory.
    void emit_local_task(int lo, int hi, std::vector<int64_t> &blocks) {
        open_task();
        loc_lo = S.c0[lo];
        loc_n_int = S.c0[hi + 1] - loc_lo;
        loc_ext.clear();
        std::vector<uint32_t> waits;
        if (upper || contributions()) {
            // every row a subtree reads from outside (backward) or contributes to (forward) lies in the below-panel row list of its root; a task may hold
            // several adjacent subtrees (a forest): the union over their roots
            for (int P = lo; P <= hi; ++P)
                if (local_root[P] == P) loc_ext.insert(loc_ext.end(), S.rows.begin() + S.rlo[P], S.rows.begin() + S.rlo[P] + S.m[P]);
            std::sort(loc_ext.begin(), loc_ext.end());
            loc_ext.erase(std::unique(loc_ext.begin(), loc_ext.end()), loc_ext.end());
            if (upper)
                for (int r : loc_ext) {
                    const uint32_t code = wait_code(task_of_panel[S.panel_of_row[r]]);
                    if (code && std::find(waits.begin(), waits.end(), code) == waits.end()) waits.push_back(code);
                }
        }
        int64_t contrib_base = 0;
        if (contributions()) {
            if ((size_t) cur >= contrib_of_task.size()) contrib_of_task.resize((size_t) cur + 1);
            contrib_base = img.n_contrib_rows;
            contrib_of_task[(size_t) cur].base = contrib_base;
            contrib_of_task[(size_t) cur].rows = loc_ext;
            img.n_contrib_rows += (int64_t) loc_ext.size();
        }
        const int n_ext = (int) loc_ext.size(), n_wait = (int) waits.size();
        stat_ext_rows += n_ext;
        img.max_local_rows = std::max(img.max_local_rows, loc_n_int + n_ext);
        auto pad4 = [](int v) { return (size_t) ((v + 3) & ~3); };
        {
            const size_t bytes = upper ? 4 * (pad4(n_ext) + pad4(n_wait)) : 4 * pad4(loc_n_int);
            const size_t off = w.record(REC_LOCAL_LOAD, loc_lo, (uint32_t) loc_n_int | (uint32_t) n_ext << 16, (uint32_t) n_wait, bytes);
            if (upper) {
                for (int e = 0; e < n_ext; ++e) w.put_int(off, (size_t) e, loc_ext[e]);
                for (int e = 0; e < n_wait; ++e) w.put_int(off, pad4(n_ext) + (size_t) e, (int32_t) waits[e]);
            } else {
                for (int i = 0; i < loc_n_int; ++i) w.put_int(off, (size_t) i, L.perm[loc_lo + i]);
            }
        }
        (void) blocks;
        emit_local_rows(lo, hi);
        for (int P = lo; P <= hi; ++P) task_of_panel[P] = cur;
        {
            if (contrib_base >= (int64_t) 1 << 31) throw std::logic_error("panel schedule: contribution buffer too large");
            const size_t off = w.record(REC_LOCAL_STORE, loc_lo, (uint32_t) loc_n_int | (upper ? 0u : (uint32_t) n_ext << 16),
                                        (uint32_t) contrib_base, upper ? 4 * pad4(loc_n_int) : 0);
            if (upper)
                for (int i = 0; i < loc_n_int; ++i) w.put_int(off, (size_t) i, L.perm[loc_lo + i]);
        }
        close_task();
    }

    void run() {
        const auto t_prep0 = std::chrono::steady_clock::now();
        const int np = S.n_panels;
        std::vector<int64_t> blocks;
        // LOCAL tasks are COMPLETE SUBTREES of the panel elimination tree (a contiguous index range,
        // the factor being postordered): everything outside such a task depends on it only through
        // its root, so fusing it adds no false dependency in either sweep, and everything it reads
        // from outside lies in the root's below-panel row list.
        std::vector<int64_t> sub_cost(np);
        std::vector<int> sub_count(np, 1), first(np);
        for (int P = 0; P < np; ++P) {
            blocks_of(P, blocks);
            int64_t cost = tri_cost(P);
            for (int64_t b : blocks) cost += block_cost(b);
            sub_cost[P] = cost;
            first[P] = P;
        }
        std::vector<int64_t> cost(sub_cost);
        for (int P = 0; P < np; ++P) {
            const int q = S.parent[P];
            if (q < 0) continue;
            sub_cost[q] += sub_cost[P];
            sub_count[q] += sub_count[P];
            first[q] = std::min(first[q], first[P]);
        }
        // measured with the local-task kernel (cfg4 / cfg5 at 128 RHS, profiles/r02x_local_sched*.log): 24 rows; up to
        // 36 / 24 is as good in f32, larger tasks cost that kernel its occupancy
        const int max_rows = opt.local_rows > 0 ? opt.local_rows : 24;
        const int max_ext = opt.local_ext > 0 ? opt.local_ext : (img.elem == 8 ? 31 : 24);  // <= 31: one loader lane per external row
        std::vector<int> root_of(np, -1);  // root of the local task a panel belongs to, or -1
        for (int P = np - 1; P >= 0; --P) {
            if (root_of[P] >= 0) continue;  // already inside a local subtree
            const bool contiguous = sub_count[P] == P - first[P] + 1;
            const int rows = S.c0[P + 1] - S.c0[first[P]];
            if (contiguous && rows <= max_rows && S.m[P] <= max_ext && rows + S.m[P] <= 255 && sub_cost[P] <= opt.local_cost)
                for (int q = first[P]; q <= P; ++q) root_of[q] = P;
        }
        // Adjacent local subtrees (consecutive in the postorder, hence independent of each other) are merged into one
        // task while the row / external-row / cost budgets allow: the fixed cost of a task (ticket, queue, slab page,
        // publish) is several times the arithmetic of a 5-row subtree.
        local_root = root_of;
        std::vector<int> group_lo(np, -1), group_hi(np, -1);
        {
            int P = 0;
            while (P < np) {
                if (root_of[P] < 0) {
                    ++P;
                    continue;
                }
                const int g_lo = P;
                int g_hi = root_of[P];
                int64_t g_cost = sub_cost[g_hi];
                std::vector<int> ext(S.rows.begin() + S.rlo[g_hi], S.rows.begin() + S.rlo[g_hi] + S.m[g_hi]), merged;
                while (g_hi + 1 < np && root_of[g_hi + 1] >= 0) {
                    const int r2 = root_of[g_hi + 1];
                    const int rows = S.c0[r2 + 1] - S.c0[g_lo];
                    merged.clear();
                    std::set_union(ext.begin(), ext.end(), S.rows.begin() + S.rlo[r2], S.rows.begin() + S.rlo[r2] + S.m[r2],
                                   std::back_inserter(merged));
                    if (rows > max_rows || (int) merged.size() > max_ext || rows + (int) merged.size() > 255 ||
                        g_cost + sub_cost[r2] > opt.local_cost)
                        break;
                    ext.swap(merged);
                    g_cost += sub_cost[r2];
                    g_hi = r2;
                }
                for (int q = g_lo; q <= g_hi; ++q) group_lo[q] = g_lo, group_hi[q] = g_hi;
                P = g_hi + 1;
            }
        }

        // Inverse super-blocks: runs of <= inv_panels consecutive panels of one supernode (outside local tasks)
        std::vector<int> sb_lo(np, -1), sb_hi(np, -1);
        if (!opt.exact_order && opt.inv_panels >= 2) {
            auto continues = [&](int P) {  // panel P continues the supernode of panel P - 1
                if (P <= 0 || root_of[P] >= 0 || root_of[P - 1] >= 0 || S.parent[P - 1] != P) return false;
                const int wd = S.width(P);
                if (S.m[P - 1] != wd + S.m[P]) return false;
                const int *R = &S.rows[(size_t) S.rlo[P - 1]];
                return R[0] == S.c0[P] && R[wd - 1] == S.c0[P + 1] - 1;
            };
            int P = 0;
            while (P < np) {
                int Q = P;
                while (Q + 1 < np && Q + 1 - P < opt.inv_panels && continues(Q + 1)) ++Q;
                if (Q > P)
                    for (int q = P; q <= Q; ++q) sb_lo[q] = P, sb_hi[q] = Q;
                P = Q + 1;
            }
        }
        std::vector<double> tinv;  // inverse of the current super-block's dense triangle, row-major nJ x nJ
        int tinv_lo = -1;
        auto prepare_inverse = [&](int lo, int hi) {
            if (tinv_lo == lo) return;
            const auto t_inv0 = std::chrono::steady_clock::now();
            tinv_lo = lo;
            const int c_lo = S.c0[lo], nJ = S.c0[hi + 1] - c_lo;
            std::vector<double> T((size_t) nJ * nJ, 0.0);
            int64_t present = 0;
            for (int j = 0; j < nJ; ++j)
                for (int e = L.colptr[c_lo + j]; e < L.colptr[c_lo + j + 1] && L.rowidx[e] < c_lo + nJ; ++e) {
                    T[(size_t) (L.rowidx[e] - c_lo) * nJ + j] = L.vals[(size_t) e];
                    present += 1;
                }
            img.n_values += present;  // the super-block's own entries are consumed through the inverse
            // Column c of the inverse: T x = e_c by forward substitution, x_i = (e_ci - sum_{q<i} T[i,q] x_q) / T[i,i] with q
            // ascending.  Columns are independent; they are computed in blocks of 32 (handed out to the host threads by
            // an atomic counter), row by row, so the inner loop runs over contiguous memory -- a top separator is one
            // super-block of a few thousand columns, n^3 / 3 multiply-adds.
            tinv.assign((size_t) nJ * nJ, 0.0);
            constexpr int kColBlock = 32;
            const int n_col_blocks = (nJ + kColBlock - 1) / kColBlock;
            std::atomic<int> next{0};
            auto worker = [&]() {
                std::vector<double> X;  // this column block of the inverse, rows [c0, nJ) x kColBlock, contiguous
                for (;;) {
                    const int cb = next.fetch_add(1);
                    if (cb >= n_col_blocks) return;
                    const int c0 = cb * kColBlock, c1 = std::min(nJ, c0 + kColBlock), wb = c1 - c0;
                    X.assign((size_t) (nJ - c0) * kColBlock, 0.0);
                    for (int i = c0; i < nJ; ++i) {
                        double *xi = &X[(size_t) (i - c0) * kColBlock];
                        const double *ti = &T[(size_t) i * nJ];
                        const int ce = std::min(wb, i - c0 + 1);  // the inverse is lower triangular
                        if (i < c1) xi[i - c0] = 1.0;
                        for (int q = c0; q < i; ++q) {
                            const double t = ti[q];
                            if (t == 0.0) continue;
                            const double *xq = &X[(size_t) (q - c0) * kColBlock];
                            const int qe = std::min(ce, q - c0 + 1);
                            for (int c = 0; c < qe; ++c) xi[c] -= t * xq[c];
                        }
                        const double d = ti[i];
                        double *out = &tinv[(size_t) i * nJ + c0];
                        for (int c = 0; c < ce; ++c) out[c] = xi[c] = xi[c] / d;
                    }
                }
            };
            const int n_threads = (int) std::max(1u, std::min(16u, std::min((unsigned) n_col_blocks, std::thread::hardware_concurrency())));
            if (n_threads <= 1 || nJ < 256) {
                worker();
            } else {
                std::vector<std::thread> pool;
                for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);
                for (std::thread &t : pool) t.join();
            }
            stat_inverse_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_inv0).count();
        };
        // panel P of the inverse super-block [lo, hi]: task(s) A (residual) then task B (inverse apply)
        auto emit_inverse_panel = [&](int P) {
            const auto t_ip0 = std::chrono::steady_clock::now();
            struct Acc { double &s; std::chrono::steady_clock::time_point t0; ~Acc() { s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } } acc_guard{stat_inverse_panel_s, t_ip0};
            const int lo = sb_lo[P], hi = sb_hi[P];
            prepare_inverse(lo, hi);
            const int c_lo = S.c0[lo], nJ = S.c0[hi + 1] - c_lo;
            blocks_of(P, blocks);
            // blocks whose partner panel lies inside the super-block are replaced by the inverse
            std::vector<int64_t> outside;
            int64_t a_cost = 0;
            for (int64_t b : blocks) {
                const int partner = upper ? S.b_tgt[b] : S.b_src[b];
                if (partner >= lo && partner <= hi) continue;
                outside.push_back(b);
                a_cost += block_cost(b);
            }
            end_store_r = true;
            if (a_cost > opt.split_cost) {
                emit_split_panel(P, outside);
            } else {
                open_task();
                emit_panel(P, outside);
                close_task();
            }
            end_store_r = false;
            const auto t_b0 = std::chrono::steady_clock::now();
            struct AccB { double &s; std::chrono::steady_clock::time_point t0; ~AccB() { s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } } accb_guard{stat_inverse_apply_s, t_b0};
            // B(P): x_P = sum over the super-block's panels p of Tinv[P, p] r_p.  Long sums are split into chunk tasks
            // that run in parallel and park partial tiles in scratch slots (as emit_split_panel does), so a hop of the
            // critical path at the top of the tree is inv_chunk blocks long, not up to inv_panels.
            const int wP = S.width(P), oP = S.c0[P] - c_lo;
            std::vector<int> srcs;  // source panels in this sweep's processing order
            if (!upper) for (int p = lo; p <= P; ++p) srcs.push_back(p);
            else for (int p = hi; p >= P; --p) srcs.push_back(p);
            auto emit_source = [&](int p) {
                const int op = S.c0[p] - c_lo, wp = S.width(p);
                if (!upper) emit_inverse_block(p, wP, atask_of_panel[p], [&](int r, int t) { return -tinv[(size_t) (oP + t) * nJ + (op + r)]; });
                // backward: sources = rows of p visited descending; unknowns = columns of P
                else emit_inverse_block(p, wP, atask_of_panel[p], [&](int r, int t) { return -tinv[(size_t) (op + wp - 1 - r) * nJ + (oP + t)]; });
            };
            const int per_chunk = opt.inv_chunk;
            int n_chunks = 1;
            if (per_chunk > 0 && (int) srcs.size() >= 2 * per_chunk) n_chunks = (int) srcs.size() / per_chunk;
            std::vector<int> slot, task;
            size_t at = 0;
            for (int c = 0; c + 1 < n_chunks; ++c) {
                task.push_back(open_task());
                emit_begin_zero(P);
                for (int q = 0; q < per_chunk; ++q) emit_source(srcs[at++]);
                slot.push_back(img.n_slots++);
                emit_end_slot(slot.back());
                close_task();
            }
            open_task();
            emit_begin_zero(P);
            for (size_t c = 0; c < slot.size(); ++c) emit_add_slot(slot[c], task[c]);
            while (at < srcs.size()) emit_source(srcs[at++]);
            emit_end_store(P, false);
            close_task();
        };

        stat_prep_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_prep0).count();
        int i = 0;
        while (i < np) {
            // [lo, hi] = the next run of panels forming one task, in this sweep's processing order
            int lo, hi;
            if (!upper) {
                lo = i;
                hi = root_of[i] >= 0 ? group_hi[i] : i;
                if (root_of[i] >= 0 && group_lo[i] != i) throw std::logic_error("panel schedule: local task is not contiguous");
            } else {
                hi = np - 1 - i;
                lo = root_of[hi] >= 0 ? group_lo[hi] : hi;
                if (root_of[hi] >= 0 && group_hi[hi] != hi) throw std::logic_error("panel schedule: local task is not contiguous");
            }
            i += hi - lo + 1;
            if (root_of[hi] >= 0) {
                const auto t_l0 = std::chrono::steady_clock::now();
                emit_local_task(lo, hi, blocks);
                stat_local_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_l0).count();
                continue;
            }
            if (sb_lo[lo] >= 0) {
                emit_inverse_panel(lo);
                continue;
            }
            blocks_of(lo, blocks);
            if (cost[lo] > opt.split_cost) {
                emit_split_panel(lo, blocks);
                continue;
            }
            open_task();
            emit_panel(lo, blocks);
            close_task();
        }
        const auto t_f0 = std::chrono::steady_clock::now();
        fill_pending_blocks();
        stat_fill_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_f0).count();
    }
};

}  // namespace

void build_sweep_image(const HostFactor &L, const PanelStructure &S, int elem_size, bool upper,
                       const ScheduleOptions &opt, SweepImage *out) {
    SweepImage &img = *out;
    img = SweepImage();
    img.upper = upper;
    img.elem = elem_size;
    img.n = L.n;
    img.page_bytes = page_bytes_for(elem_size);
    img.slab.reserve((size_t) ((double) L.nnz() * elem_size * 1.3) + (size_t) S.n_panels * 160 + 4096);
    Builder b(L, S, opt, img);
    const auto t_run0 = std::chrono::steady_clock::now();
    b.run();
    const double run_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run0).count();
    if (std::getenv("CHB_SCHED_STATS"))
        fprintf(stderr, "[sched %s] build %.2f s: preparation %.2f s, local tasks %.2f s, inverse super-block panels %.2f s (triangle inverses %.2f s, inverse-apply tasks %.2f s), block payloads %.2f s\n",
                upper ? "bwd" : "fwd", run_s, b.stat_prep_s, b.stat_local_s, b.stat_inverse_panel_s, b.stat_inverse_s, b.stat_inverse_apply_s, b.stat_fill_s);
    if (std::getenv("CHB_SCHED_STATS"))
        fprintf(stderr, "[sched %s] non-local blocks %lld (rows %lld), of which from local producers %lld (rows %lld, values %lld); ext rows of local tasks %lld\n",
                upper ? "bwd" : "fwd", (long long) b.stat_blocks_nonlocal, (long long) b.stat_rows_nonlocal, (long long) b.stat_blocks_from_local,
                (long long) b.stat_rows_from_local, (long long) b.stat_vals_from_local, (long long) b.stat_ext_rows);

    // ticket order: by dependency level, larger slabs first inside a level
    const int nt = (int) img.tasks.size();
    uint32_t max_level = 0;
    for (const TaskDesc &t : img.tasks) max_level = std::max(max_level, t.level);
    img.n_levels = nt ? (int) max_level + 1 : 0;
    img.order.resize(nt);
    std::iota(img.order.begin(), img.order.end(), 0);
    std::stable_sort(img.order.begin(), img.order.end(), [&](int a, int c) {
        const TaskDesc &x = img.tasks[a], &y = img.tasks[c];
        if (x.level != y.level) return x.level < y.level;
        return x.bytes > y.bytes;
    });
}

}  // namespace chb
