// Supernodal "front" schedule of the factor for the sm_100a front sweep kernels.  See front_schedule.h.
#include "front_schedule.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace chb {

namespace {

inline size_t round16(size_t b) { return (b + 15) & ~size_t(15); }
inline size_t pad4(size_t v) { return (v + 3) & ~size_t(3); }

// Appends records to the image; a record never straddles a page of its task's slab.
class Writer {
public:
    Writer(std::vector<uint8_t> &buf, int page, int elem) : m_buf(buf), m_page((size_t) page), m_elem(elem) {}
    void begin_task() { m_task_start = m_buf.size(); }
    size_t size() const { return m_buf.size(); }
    // Reserves a zero-filled record; returns the byte offset of its payload.  Every page starts with a DIRECTORY record
    // listing the page's staged (FR_LOADX) records, so that the loader warp jumps to them instead of walking every record.
    size_t record(uint32_t type, uint32_t a, uint32_t b, uint32_t c, size_t payload_bytes) {
        const size_t len = 16 + round16(payload_bytes);
        if (len > m_page - 16) throw std::logic_error("front schedule: record larger than a page");
        size_t in_page = (m_buf.size() - m_task_start) % m_page;
        if (in_page != 0 && in_page + len > m_page) {
            header(FR_NOP, 0, 0, 0, m_page - in_page);
            in_page = 0;
        }
        if (in_page == 0) start_page();
        const size_t off = m_buf.size();
        header(type, a, b, c, len);
        if (type == FR_LOADX) {
            uint32_t h0;
            std::memcpy(&h0, &m_buf[m_dir], 4);
            const uint32_t cnt = h0 >> 24;
            if (cnt < 6) {
                const uint16_t units = (uint16_t) ((off - m_dir) / 16);
                std::memcpy(&m_buf[m_dir + 4 + 2 * cnt], &units, 2);
                h0 += 1u << 24;
            } else {
                h0 = (h0 & 0x00ffffffu) | 7u << 24;  // more than six: the loader walks the page
            }
            std::memcpy(&m_buf[m_dir], &h0, 4);
        }
        return off + 16;
    }
    // bytes a record may still take in the current page of the current task
    size_t room_in_page() const {
        const size_t in_page = (m_buf.size() - m_task_start) % m_page;
        return in_page == 0 ? m_page - 16 : m_page - in_page;
    }
    size_t page_overhead() const { return 16; }
    void pad_page() {
        const size_t in_page = (m_buf.size() - m_task_start) % m_page;
        if (in_page) header(FR_NOP, 0, 0, 0, m_page - in_page);
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
    void start_page() {
        m_dir = m_buf.size();
        header(FR_DIR, 0, 0, 0, 16);
    }
    void header(uint32_t type, uint32_t a, uint32_t b, uint32_t c, size_t len) {
        const size_t off = m_buf.size();
        m_buf.resize(off + len);  // zero-filled
        const uint32_t h[4] = {type | (uint32_t) (len / 16) << 8, a, b, c};
        std::memcpy(&m_buf[off], h, 16);
    }
    std::vector<uint8_t> &m_buf;
    size_t m_page;
    int m_elem;
    size_t m_task_start = 0;
    size_t m_dir = 0;  // offset of the current page's directory record
};

int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

}  // namespace

FrontOptions front_options_from_env(int elem) {
    FrontOptions o;
    o.warps = env_int("CHB_FRONT_WARPS", 7) == 4 ? 4 : 7;
    // the row ring of a CTA holds (at 128 right-hand sides) about 174 / 120 single-precision rows (7 / 4 warps), half
    // as many double-precision ones: a local task (rows + external rows) and a panel must fit with a chunk to spare
    const int cap = (o.warps == 7 ? 168 : 112) / (elem == 8 ? 2 : 1);
    o.local_rows = std::max(1, std::min(240, env_int("CHB_FRONT_LOCAL_ROWS", cap * 2 / 3)));
    o.local_ext = std::max(0, std::min(128, env_int("CHB_FRONT_LOCAL_EXT", cap / 3)));
    const int full = o.warps * front_tile_rows(elem);
    o.panel = std::max(front_tile_rows(elem), std::min(full, env_int("CHB_FRONT_PANEL", full)));
    o.fuse_steps = std::max(0, env_int("CHB_FRONT_FUSE_STEPS", o.fuse_steps));
    o.split_chunks = std::max(1, env_int("CHB_FRONT_SPLIT_CHUNKS", o.split_chunks));
    o.own_chunks = std::max(1, env_int("CHB_FRONT_OWN_CHUNKS", o.own_chunks));
    o.order_blocks = std::max(0, env_int("CHB_FRONT_ORDER_BLOCKS", o.order_blocks));
    return o;
}

void build_front_structure(const HostFactor &L, int elem, const FrontOptions &opt, FrontStructure *out) {
    (void) elem;
    FrontStructure &S = *out;
    S = FrontStructure();
    const int n = L.n;
    const std::vector<int> &cp = L.colptr, &ri = L.rowidx;
    S.n = n;
    S.sn_of.resize(n);
    // Maximal supernodes: column j + 1 continues the supernode of j when it is j's etree parent and its structure is
    // j's minus the diagonal (then the two columns have the same below-rows: a dense block column without fill).
    for (int j = 0; j < n; ++j) {
        const bool joins = j > 0 && cp[j] - cp[j - 1] > 1 && ri[cp[j - 1] + 1] == j && cp[j + 1] - cp[j] == cp[j] - cp[j - 1] - 1;
        if (!joins) S.c0.push_back(j);
        S.sn_of[j] = (int) S.c0.size() - 1;
    }
    S.ns = (int) S.c0.size();
    S.c0.push_back(n);
    const int ns = S.ns;
    S.parent.assign(ns, -1);
    auto below = [&](int s) { return cp[S.c0[s] + 1] - cp[S.c0[s]] - S.width(s); };      // |R_s|
    auto below_rows = [&](int s) { return &ri[(size_t) cp[S.c0[s]] + (size_t) S.width(s)]; };  // R_s (sorted)
    for (int s = 0; s < ns; ++s)
        if (below(s) > 0) S.parent[s] = S.sn_of[below_rows(s)[0]];

    // LOCAL tasks: complete subtrees (contiguous index ranges, the factor being postordered) of at most local_rows
    // rows whose root has at most local_ext below-rows; adjacent ones are merged while the budgets allow.
    std::vector<int> sub_rows(ns), sub_count(ns, 1), first(ns);
    for (int s = 0; s < ns; ++s) sub_rows[s] = S.width(s), first[s] = s;
    for (int s = 0; s < ns; ++s) {
        const int q = S.parent[s];
        if (q < 0) continue;
        if (q <= s) throw std::logic_error("front schedule: elimination tree is not topologically ordered");
        sub_rows[q] += sub_rows[s];
        sub_count[q] += sub_count[s];
        first[q] = std::min(first[q], first[s]);
    }
    std::vector<int> root_of(ns, -1);
    for (int s = ns - 1; s >= 0; --s) {
        if (root_of[s] >= 0) continue;
        const bool contiguous = sub_count[s] == s - first[s] + 1;
        if (contiguous && sub_rows[s] <= opt.local_rows && below(s) <= opt.local_ext && sub_rows[s] + below(s) <= 255)
            for (int q = first[s]; q <= s; ++q) root_of[q] = s;
    }
    S.local_of.assign(ns, -1);
    int s = 0;
    std::vector<int> merged;
    while (s < ns) {
        if (root_of[s] < 0) {
            ++s;
            continue;
        }
        FrontStructure::Local loc;
        loc.sn_lo = s;
        loc.sn_hi = root_of[s];
        loc.roots.push_back(loc.sn_hi);
        loc.ext.assign(below_rows(loc.sn_hi), below_rows(loc.sn_hi) + below(loc.sn_hi));
        while (loc.sn_hi + 1 < ns && root_of[loc.sn_hi + 1] >= 0) {
            const int r2 = root_of[loc.sn_hi + 1];
            const int rows = S.c0[r2 + 1] - S.c0[loc.sn_lo];
            merged.clear();
            std::set_union(loc.ext.begin(), loc.ext.end(), below_rows(r2), below_rows(r2) + below(r2), std::back_inserter(merged));
            if (rows > opt.local_rows || (int) merged.size() > opt.local_ext || rows + (int) merged.size() > 255) break;
            loc.ext.swap(merged);
            loc.sn_hi = r2;
            loc.roots.push_back(r2);
        }
        loc.c_lo = S.c0[loc.sn_lo];
        loc.c_hi = S.c0[loc.sn_hi + 1];
        for (int q = loc.sn_lo; q <= loc.sn_hi; ++q) S.local_of[q] = (int) S.locals.size();
        s = loc.sn_hi + 1;
        S.locals.push_back(std::move(loc));
    }

    // Subtree blocks of the ticket order: maximal subtrees of at most n / order_blocks rows, numbered in postorder;
    // everything above them is the last block.
    S.block_of.assign(ns, 0);
    S.n_blocks = 1;
    if (opt.order_blocks > 1) {
        const int thr = std::max(1, n / opt.order_blocks);
        int nb = 0;
        std::vector<int> blk(ns, -1);
        for (int q = 0; q < ns; ++q) {
            const int p = S.parent[q];
            if (sub_rows[q] <= thr && (p < 0 || sub_rows[p] > thr)) blk[q] = nb++;
        }
        for (int q = ns - 1; q >= 0; --q) {
            const int p = S.parent[q];
            if (blk[q] < 0 && p >= 0 && sub_rows[q] <= thr) blk[q] = blk[p];
        }
        for (int q = 0; q < ns; ++q) S.block_of[q] = blk[q] < 0 ? nb : blk[q];
        S.n_blocks = nb + 1;
    }
}

namespace {

struct FrontBuilder {
    const HostFactor &L;
    const FrontStructure &S;
    const FrontOptions &opt;
    FrontImage &img;
    Writer w;
    const bool upper;
    const int TR, P, es, NW;
    const std::vector<int> &cp, &ri;

    int cur = -1;
    uint32_t cur_level = 0;
    int cur_sj = 0;
    std::vector<int> block_of_task;
    std::vector<std::vector<int>> done_of;        // per supernode: tasks a consumer of its outputs waits for
    std::vector<std::vector<int>> panel_task;     // per dense supernode: the task that publishes each panel's unknowns
    std::vector<int> local_task;                  // per local: task id
    std::vector<std::vector<int>> children;       // per supernode
    std::vector<int> dense_row;                   // n: index into contrib, or -1
    std::vector<std::vector<int32_t>> contrib;    // forward: per dense row, the W rows addressed to it

    FrontBuilder(const HostFactor &L_, const FrontStructure &S_, const FrontOptions &o, FrontImage &i)
        : L(L_), S(S_), opt(o), img(i), w(i.slab, i.page_bytes, i.elem), upper(i.upper), TR(i.tile_rows), P(o.panel),
          es(i.elem), NW(o.warps), cp(L_.colptr), ri(L_.rowidx) {}

    int width(int s) const { return S.width(s); }
    int below(int s) const { return cp[S.c0[s] + 1] - cp[S.c0[s]] - width(s); }
    const int *below_rows(int s) const { return &ri[(size_t) cp[S.c0[s]] + (size_t) width(s)]; }
    // L[row, col] for a row inside col's supernode (row >= col)
    double in_sn(int col, int row) const { return L.vals[(size_t) cp[col] + (size_t) (row - col)]; }
    // L[R_s[q], col] for col in supernode s
    double below_val(int s, int col, int q) const { return L.vals[(size_t) cp[col] + (size_t) (S.c0[s + 1] - col) + (size_t) q]; }

    // -- tasks ----------------------------------------------------------------------------------
    int open_task(int sn) {
        cur = (int) img.tasks.size();
        img.tasks.push_back(TaskDesc{(uint64_t) w.size(), 0u, 0u});
        block_of_task.push_back(S.block_of[sn]);
        w.begin_task();
        cur_level = 0;
        cur_sj = 0;
        return cur;
    }
    void close_task() {
        const uint64_t bytes = (uint64_t) w.size() - img.tasks[cur].offset;
        if (bytes >= (1ull << 31)) throw std::logic_error("front schedule: task slab too large");
        img.tasks[cur].bytes = (uint32_t) bytes;
        img.tasks[cur].level = cur_level;
        cur = -1;
    }
    void depend(int task) {
        if (task < 0 || task >= cur) throw std::logic_error("front schedule: dependency on an unscheduled task");
        cur_level = std::max(cur_level, img.tasks[task].level + 1);
    }

    // -- records --------------------------------------------------------------------------------
    // Returns the staged index of the record.
    int emit_loadx(int contig_buf, int row0, int n_contig, int gather_kind, const int *gather, int n_gather, int n_extra,
                   std::vector<int> waits) {
        std::sort(waits.begin(), waits.end());
        waits.erase(std::unique(waits.begin(), waits.end()), waits.end());
        for (int t : waits) depend(t);
        const int total = n_contig + n_gather + n_extra;
        if (total > 1023 || n_contig > 1023 || n_gather > 1023 || n_extra > 1023 || waits.size() > 0xffff)
            throw std::logic_error("front schedule: staged record too large");
        img.max_stage_rows = std::max(img.max_stage_rows, total);
        const size_t off = w.record(FR_LOADX, (uint32_t) n_contig | (uint32_t) n_gather << 10 | (uint32_t) n_extra << 20, (uint32_t) row0,
                                    (uint32_t) contig_buf | (uint32_t) gather_kind << 4 | (uint32_t) waits.size() << 8,
                                    4 * (pad4(waits.size()) + pad4((size_t) n_gather)));
        for (size_t i = 0; i < waits.size(); ++i) w.put_int(off, i, waits[i]);
        for (int i = 0; i < n_gather; ++i) w.put_int(off, pad4(waits.size()) + (size_t) i, gather[i]);
        return cur_sj++;
    }
    // Reserves an MMA record; returns the payload offset of tile 0's values ([n_src][TR] per tile).
    size_t emit_mma(int tile0, int n_tiles, int n_src, bool zero_init, int sj, int src_row) {
        if (n_tiles < 1 || tile0 + n_tiles > NW || n_src < 1 || n_src > kFrontChunk || sj > 0xffff || src_row > 0xffff)
            throw std::logic_error("front schedule: bad MMA record");
        img.n_mma += 1;
        img.n_stored_values += (int64_t) n_tiles * n_src * TR;
        return w.record(FR_MMA, (uint32_t) tile0 | (uint32_t) n_tiles << 8 | (uint32_t) n_src << 16 | (zero_init ? 1u << 24 : 0u),
                        (uint32_t) sj | (uint32_t) src_row << 16, 0, (size_t) n_tiles * (size_t) n_src * (size_t) TR * (size_t) es);
    }
    void put_a(size_t off, int tile, int n_src, int s, int i, double v) { w.put_value(off, ((size_t) tile * n_src + s) * TR + i, v); }
    void emit_epi(int n_tiles, uint32_t mode, int n_rows, int sj, int ring_row, int global_row, const int *perm) {
        const size_t off = w.record(FR_EPI, (uint32_t) 0 | (uint32_t) n_tiles << 8 | mode << 16, (uint32_t) n_rows | (uint32_t) sj << 16,
                                    (uint32_t) global_row, 16 + ((mode & FE_DST) ? 4 * pad4((size_t) n_rows) : 0));
        w.put_int(off, 0, ring_row);
        if (mode & FE_DST)
            for (int i = 0; i < n_rows; ++i) w.put_int(off, 4 + (size_t) i, perm[i]);
    }
    void emit_bar() { w.record(FR_BAR, 0, 0, 0, 0); }
    void emit_rel(int sj) { w.record(FR_REL, (uint32_t) sj, 0, 0, 0); }
    // ring[sj][ring_row + r] -= W[w_row] for the given (r, w_row) pairs; split over records that fit a page
    void emit_asm(int sj, int ring_row, std::vector<std::pair<int, int32_t>> &entries) {
        if (entries.empty()) return;
        // warp w owns rows r == w (mod n_warps): sort by (r mod n_warps, r), stable in W row order
        const int nw = NW;
        std::stable_sort(entries.begin(), entries.end(), [nw](const std::pair<int, int32_t> &x, const std::pair<int, int32_t> &y) {
            return (x.first % nw) != (y.first % nw) ? (x.first % nw) < (y.first % nw) : x.first < y.first;
        });
        const size_t cap = ((size_t) img.page_bytes - 16 - 16 - 48) / 8;
        // a record holds complete rows only (a row's entries must be applied by one warp in one go is not required, but
        // keeping per-warp segments sorted is): simply cut the per-warp lists proportionally
        std::vector<size_t> seg_lo((size_t) NW + 1, 0);
        for (const auto &e : entries) seg_lo[(size_t) (e.first % NW) + 1] += 1;
        for (int i = 0; i < NW; ++i) seg_lo[(size_t) i + 1] += seg_lo[(size_t) i];
        std::vector<size_t> at(seg_lo.begin(), seg_lo.end() - 1);
        size_t remaining = entries.size();
        while (remaining > 0) {
            const size_t per_warp = std::max<size_t>(1, cap / NW);
            size_t take[kFrontMaxWarps], total = 0;
            for (int i = 0; i < NW; ++i) {
                take[i] = std::min(per_warp, seg_lo[(size_t) i + 1] - at[(size_t) i]);
                total += take[i];
            }
            const size_t off = w.record(FR_ASM, (uint32_t) sj | (uint32_t) ring_row << 16, (uint32_t) total, 0, 48 + 8 * total);
            size_t pos = 0;
            for (int i = 0; i < NW; ++i) {
                w.put_int(off, (size_t) i, (int32_t) pos);
                for (size_t e = 0; e < take[i]; ++e) {
                    const auto &en = entries[at[(size_t) i] + e];
                    w.put_int(off + 48, 2 * (pos + e), en.first);
                    w.put_int(off + 48, 2 * (pos + e) + 1, en.second);
                }
                pos += take[i];
                at[(size_t) i] += take[i];
            }
            w.put_int(off, NW, (int32_t) pos);
            remaining -= total;
            img.n_asm_entries += (int64_t) total;
        }
    }

    // Inverse of the dense lower triangle L[p0 .. p0 + pw, p0 .. p0 + pw) of one supernode, row-major pw x pw (double).
    void invert_diag(int p0, int pw, std::vector<double> &inv) const {
        inv.assign((size_t) pw * pw, 0.0);
        for (int c = 0; c < pw; ++c) {  // column c of the inverse: T x = e_c by forward substitution
            for (int i = c; i < pw; ++i) {
                double acc = i == c ? 1.0 : 0.0;
                for (int q = c; q < i; ++q) acc -= in_sn(p0 + q, p0 + i) * inv[(size_t) q * pw + c];
                inv[(size_t) i * pw + c] = acc / in_sn(p0 + i, p0 + i);
            }
        }
    }
    // y = inv(T) v (forward) / x = inv(T)' v (backward) for the pw rows staged at ring row `ring_row` of record sj
    void emit_inverse(int p0, int pw, int sj, int ring_row) {
        std::vector<double> inv;
        invert_diag(p0, pw, inv);
        img.n_values += (int64_t) pw * (pw + 1) / 2;
        const int nt = (pw + TR - 1) / TR;
        // forward: tile t needs sources < (t + 1) TR, chunks ascending; backward: sources >= t TR, chunks descending.
        // Either way the first chunk visited covers every tile (it carries zero_init) and later ones a sub-range.
        const int n_chunks = (pw + kFrontChunk - 1) / kFrontChunk;
        for (int ci = 0; ci < n_chunks; ++ci) {
            const int s0 = (upper ? n_chunks - 1 - ci : ci) * kFrontChunk;
            const int ns = std::min(kFrontChunk, pw - s0);
            const int t_lo = upper ? 0 : s0 / TR;
            const int t_hi = upper ? std::min(nt, (s0 + ns + TR - 1) / TR) : nt;
            const size_t off = emit_mma(t_lo, t_hi - t_lo, ns, ci == 0, sj, ring_row + s0);
            for (int t = t_lo; t < t_hi; ++t)
                for (int s = 0; s < ns; ++s)
                    for (int i = 0; i < TR; ++i) {
                        const int r = t * TR + i, c = s0 + s;
                        if (r >= pw) continue;
                        const double v = upper ? (c >= r ? inv[(size_t) c * pw + r] : 0.0) : (c <= r ? inv[(size_t) r * pw + c] : 0.0);
                        put_a(off, t - t_lo, ns, s, i, v);
                    }
        }
    }

    // One source chunk of an accumulation: <= 32 rows of Y.
    struct Chunk {
        bool gather = false;  // rows listed (backward below-rows) / contiguous
        int row0 = 0;         // contiguous: first row;  gather: index into the supernode's below-row list
        int n = 0;
        int wait = -1;        // task publishing these rows (-1: covered by an earlier wait)
        int kind = 0;         // which factor block the values come from (see fill_block)
    };
    // values of the block (target rows x chunk rows), kinds:
    //   0 forward, target = rows [t_row0, +) of supernode s's columns, source = columns [row0, row0 + n) of s (before them)
    //   1 forward, target = below-rows R_s[t_row0 ..], source = columns [row0, row0 + n) of s
    //   2 backward, target = columns [t_row0, +) of s, source = rows [row0, row0 + n) of s (later columns' rows)
    //   3 backward, target = columns [t_row0, +) of s, source = below-rows R_s[row0 .. row0 + n)
    void fill_block(size_t off, int s, int kind, int t_row0, int t_rows, const Chunk &c) {
        const int nt = (t_rows + TR - 1) / TR;
        for (int t = 0; t < nt; ++t)
            for (int i = 0; i < TR; ++i) {
                const int tr = t * TR + i;
                if (tr >= t_rows) break;
                for (int q = 0; q < c.n; ++q) {
                    double v;
                    switch (kind) {
                    case 0: v = in_sn(c.row0 + q, t_row0 + tr); break;
                    case 1: v = below_val(s, c.row0 + q, t_row0 + tr); break;
                    case 2: v = in_sn(t_row0 + tr, c.row0 + q); break;
                    default: v = below_val(s, t_row0 + tr, c.row0 + q); break;
                    }
                    put_a(off, t, c.n, q, i, v);
                }
            }
        img.n_values += (int64_t) t_rows * c.n;
    }
    // acc[tiles of the target] (+)= sum over chunks; every chunk is staged, multiplied and released
    void emit_accumulate(int s, int t_row0, int t_rows, const std::vector<Chunk> &chunks, bool zero_first) {
        const int nt = (t_rows + TR - 1) / TR;
        bool first = zero_first;
        for (const Chunk &c : chunks) {
            std::vector<int> waits;
            if (c.wait >= 0) waits.push_back(c.wait);
            int sj;
            if (c.gather) sj = emit_loadx(0, 0, 0, 1, below_rows(s) + c.row0, c.n, 0, waits);
            else sj = emit_loadx(1, c.row0, c.n, 0, nullptr, 0, 0, waits);
            for (int o = 0; o < c.n; o += kFrontChunk) {
                Chunk sub = c;
                sub.row0 = c.row0 + o;
                sub.n = std::min(kFrontChunk, c.n - o);
                const size_t off = emit_mma(0, nt, sub.n, first, sj, o);
                fill_block(off, s, c.kind, t_row0, t_rows, sub);
                first = false;
            }
            emit_rel(sj);
        }
    }

    // -- local tasks ------------------------------------------------------------------------------
    void emit_local(int li) {
        const FrontStructure::Local &loc = S.locals[(size_t) li];
        const int n_int = loc.c_hi - loc.c_lo, n_ext = (int) loc.ext.size();
        open_task(loc.sn_lo);
        img.n_local_tasks += 1;
        auto local_index = [&](int row) {
            if (row < loc.c_hi) return row - loc.c_lo;
            const auto it = std::lower_bound(loc.ext.begin(), loc.ext.end(), row);
            if (it == loc.ext.end() || *it != row) throw std::logic_error("front schedule: local task touches a row it did not load");
            return n_int + (int) (it - loc.ext.begin());
        };
        int sj;
        int64_t w0 = 0;
        if (!upper) {
            sj = emit_loadx(0, 0, 0, 2, &L.perm[(size_t) loc.c_lo], n_int, n_ext, {});
            w0 = img.n_w_rows;
            img.n_w_rows += n_ext;
            for (int e = 0; e < n_ext; ++e) contrib[(size_t) dense_row[(size_t) loc.ext[(size_t) e]]].push_back((int32_t) (w0 + e));
        } else {
            std::vector<int> waits;
            for (int r : loc.roots)
                if (below(r) > 0) waits.push_back(owner_task(below_rows(r)[0]));
            sj = emit_loadx(1, loc.c_lo, n_int, 1, loc.ext.data(), n_ext, 0, waits);
        }
        // sparse rows: forward row i of L by ascending column, backward column j of L by descending row; entries in
        // groups of 4 (zero padding), reciprocal pivots
        struct Entry {
            int idx;
            double v;
        };
        const int n_out = upper ? 0 : n_ext;
        std::vector<std::vector<Entry>> rows((size_t) (n_int + n_out));
        std::vector<double> diag((size_t) (n_int + n_out), 1.0);
        for (int j = 0; j < n_int; ++j) {
            const int c = loc.c_lo + j;
            diag[(size_t) j] = L.vals[(size_t) cp[c]];
            if (!upper) {
                for (int e = cp[c] + 1; e < cp[c + 1]; ++e) rows[(size_t) local_index(ri[e])].push_back(Entry{j, L.vals[(size_t) e]});
            } else {
                for (int e = cp[c + 1] - 1; e > cp[c]; --e) rows[(size_t) j].push_back(Entry{local_index(ri[e]), L.vals[(size_t) e]});
            }
            img.n_values += cp[c + 1] - cp[c];
        }
        const int n_rows_total = n_int + n_out;
        // Dealing the rows to the compute warps (one warp substitutes one row over all RHS columns).
        //   * SUBTREES: the task's rows form a forest (its elimination trees); subtrees of about 1/n_warps of the work are
        //     given whole to one warp, which substitutes their rows one after the other -- a row depends only on rows of
        //     its own subtree (forward: descendants, done before; backward: ancestors, see below), so this phase needs no
        //     synchronisation at all;
        //   * what is left -- the few rows ABOVE those subtrees, and (forward) the contribution rows, which collect
        //     from several subtrees -- is done level by level, the rows of a level dealt to the warps, a barrier between
        //     levels.  Forward: subtrees, barrier, upper levels.  Backward: upper levels, then the subtrees.
        std::vector<int> lparent((size_t) n_int, -1);
        for (int j = 0; j < n_int; ++j) {
            const int c = loc.c_lo + j;
            if (cp[c + 1] - cp[c] > 1 && ri[cp[c] + 1] < loc.c_hi) lparent[(size_t) j] = ri[cp[c] + 1] - loc.c_lo;
        }
        auto cost_of = [&](int r) { return (int64_t) rows[(size_t) r].size() / 4 + 2; };
        std::vector<int64_t> sub_w((size_t) n_int);
        int64_t total_w = 0;
        for (int j = 0; j < n_int; ++j) sub_w[(size_t) j] = cost_of(j), total_w += cost_of(j);
        for (int j = 0; j < n_int; ++j)
            if (lparent[(size_t) j] >= 0) sub_w[(size_t) lparent[(size_t) j]] += sub_w[(size_t) j];
        const int64_t target = std::max<int64_t>(total_w / NW + 1, 12);
        std::vector<int> group((size_t) n_int, -1);  // warp subtree of a row, or -1: upper row
        std::vector<int64_t> group_w;
        for (int j = n_int - 1; j >= 0; --j) {
            const int q = lparent[(size_t) j];
            if (q >= 0 && group[(size_t) q] >= 0) {
                group[(size_t) j] = group[(size_t) q];
            } else if (sub_w[(size_t) j] <= target) {
                group[(size_t) j] = (int) group_w.size();
                group_w.push_back(sub_w[(size_t) j]);
            }
        }
        std::vector<int> warp_of_group(group_w.size());
        {
            std::vector<int> by_w(group_w.size());
            std::iota(by_w.begin(), by_w.end(), 0);
            std::stable_sort(by_w.begin(), by_w.end(), [&](int x, int y) { return group_w[(size_t) x] > group_w[(size_t) y]; });
            int64_t load[kFrontMaxWarps] = {0};
            for (int gi : by_w) {
                int best = 0;
                for (int wi = 1; wi < NW; ++wi)
                    if (load[wi] < load[best]) best = wi;
                warp_of_group[(size_t) gi] = best;
                load[best] += group_w[(size_t) gi];
            }
        }
        // What is emitted: "virtual rows" {target row, entries, pivot, accumulate from zero}.  A subtree row or a
        // contribution row is one virtual row.  An UPPER row is split in two: (A) its entries from outside the upper set
        // -- subtree rows forward, external rows backward, all final when the upper phase starts -- applied with pivot 1
        // by any warp (one parallel phase), and (B) its entries from other upper rows plus the pivot, which form the
        // short sequential chain at the top of the task.
        struct VRow {
            int row;
            std::vector<Entry> ent;
            double pivot;
            bool from_zero;
        };
        std::vector<VRow> vrows;
        std::vector<int> v_sub((size_t) n_int, -1), v_a((size_t) n_int, -1), v_b((size_t) n_int, -1), v_contrib;
        auto is_upper = [&](int r) { return r < n_int && group[(size_t) r] < 0; };
        for (int r = 0; r < n_rows_total; ++r) {
            if (r >= n_int) {
                v_contrib.push_back((int) vrows.size());
                vrows.push_back(VRow{r, rows[(size_t) r], -1.0, true});  // stored as +sum: consumers subtract
            } else if (!is_upper(r)) {
                v_sub[(size_t) r] = (int) vrows.size();
                vrows.push_back(VRow{r, rows[(size_t) r], 1.0 / diag[(size_t) r], false});
            } else {
                VRow a{r, {}, 1.0, false}, b{r, {}, 1.0 / diag[(size_t) r], false};
                for (const Entry &e : rows[(size_t) r]) (is_upper(e.idx) ? b.ent : a.ent).push_back(e);
                if (!a.ent.empty()) {
                    v_a[(size_t) r] = (int) vrows.size();
                    vrows.push_back(std::move(a));
                }
                v_b[(size_t) r] = (int) vrows.size();
                vrows.push_back(std::move(b));
            }
        }
        auto vcost = [&](int v) { return (int64_t) vrows[(size_t) v].ent.size() / 4 + 2; };
        // levels among the (B) parts of the upper rows
        std::vector<int> level((size_t) n_int, -1);
        int n_lv = 0;
        for (int q = 0; q < n_int; ++q) {
            const int r = upper ? n_int - 1 - q : q;  // dependencies first
            if (!is_upper(r)) continue;
            int lv = 0;
            for (const Entry &e : vrows[(size_t) v_b[(size_t) r]].ent) lv = std::max(lv, level[(size_t) e.idx] + 1);
            level[(size_t) r] = lv;
            n_lv = std::max(n_lv, lv + 1);
        }
        // (B) level by level (a barrier each, ~8 cost units) or one after the other on one warp?
        bool upper_serial = false;
        {
            std::vector<int64_t> lv_max((size_t) n_lv, 0);
            int64_t serial = 0;
            for (int r = 0; r < n_int; ++r)
                if (level[(size_t) r] >= 0) {
                    serial += vcost(v_b[(size_t) r]);
                    lv_max[(size_t) level[(size_t) r]] = std::max(lv_max[(size_t) level[(size_t) r]], vcost(v_b[(size_t) r]));
                }
            int64_t leveled = 0;
            for (int64_t v : lv_max) leveled += v + 8;
            upper_serial = serial <= leveled;
        }
        img.n_local_levels += (n_lv > 0 ? 1 : 0) + (upper_serial ? (n_lv > 0 ? 1 : 0) : n_lv) + 1 + (n_out > 0 ? 1 : 0);
        img.n_local_rows += (int64_t) vrows.size();
        const size_t head_bytes = 2 * (size_t) es;  // {u32 meta, pivot}: pivot at byte es, indices from byte 2 es
        auto row_bytes = [&](int v) {
            const size_t n4 = (vrows[(size_t) v].ent.size() + 3) / 4;
            return round16(head_bytes + 4 * n4) + 4 * n4 * (size_t) es;
        };
        std::vector<std::vector<int>> of_warp((size_t) NW);  // virtual rows of the phase being emitted, per warp, in order
        // one phase: every warp substitutes its list in order; records as large as the pages allow; then a barrier
        int phase_kind = 0;
        auto emit_phase = [&]() {
            size_t at_w[kFrontMaxWarps] = {0};
            size_t left = 0;
            for (const auto &v : of_warp) left += v.size();
            if (left == 0) return;
            {
                int64_t worst_g = 0, worst_r = 0;
                for (const auto &v : of_warp) {
                    int64_t g = 0;
                    for (int x : v) g += (int64_t) (vrows[(size_t) x].ent.size() + 3) / 4;
                    if (g + (int64_t) v.size() > worst_g + worst_r) worst_g = g, worst_r = (int64_t) v.size();
                }
                img.local_phase_groups[phase_kind] += worst_g;
                img.local_phase_rows[phase_kind] += worst_r;
            }
            while (left > 0) {
                size_t room = w.room_in_page();
                if (room < 16 + 48 + 256) room = (size_t) img.page_bytes - w.page_overhead();
                room -= 16 + 48;
                size_t take[kFrontMaxWarps] = {0}, bytes = 0;
                // rows are taken evenly from the warps' lists, keeping each list's order
                bool progress = true;
                while (progress) {
                    progress = false;
                    for (int wi = 0; wi < NW; ++wi) {
                        const size_t q = at_w[wi] + take[wi];
                        if (q >= of_warp[(size_t) wi].size()) continue;
                        const size_t rb = row_bytes(of_warp[(size_t) wi][q]);
                        if (bytes + rb > room) continue;
                        bytes += rb;
                        take[wi] += 1;
                        progress = true;
                    }
                }
                if (bytes == 0) {
                    if (room + 16 + 48 < (size_t) img.page_bytes - w.page_overhead()) {  // retry on a fresh page
                        w.pad_page();
                        continue;
                    }
                    throw std::logic_error("front schedule: a local row does not fit a page");
                }
                const size_t off = w.record(FR_LROWS, 0, (uint32_t) sj, 0, 48 + bytes);
                size_t at = off + 48;
                for (int wi = 0; wi < NW; ++wi) {
                    w.put_int(off, (size_t) wi, (int32_t) (at - off));
                    for (size_t q = 0; q < take[wi]; ++q) {
                        const VRow &vr = vrows[(size_t) of_warp[(size_t) wi][at_w[wi] + q]];
                        const std::vector<Entry> &ent = vr.ent;
                        const size_t n4 = (ent.size() + 3) / 4;
                        if (n4 > 255) throw std::logic_error("front schedule: local row too long");
                        w.put_int(at, 0, (int32_t) ((uint32_t) vr.row | (uint32_t) n4 << 8 | (vr.from_zero ? 1u << 16 : 0u)));
                        w.put_value(at + (size_t) es, 0, vr.pivot);
                        uint8_t idx[1024];
                        // padding entries (value 0) name the row's first source: a finished, finite row
                        for (size_t i = 0; i < 4 * n4; ++i) idx[i] = (uint8_t) (i < ent.size() ? ent[i].idx : ent[0].idx);
                        w.put_bytes(at + head_bytes, idx, 4 * n4);
                        const size_t voff = at + round16(head_bytes + 4 * n4);
                        for (size_t i = 0; i < ent.size(); ++i) w.put_value(voff, i, ent[i].v);
                        at = voff + 4 * n4 * (size_t) es;
                        img.n_stored_values += (int64_t) (4 * n4) + 1;
                    }
                    at_w[wi] += take[wi];
                    left -= take[wi];
                }
                w.put_int(off, (size_t) NW, (int32_t) (at - off));
            }
            emit_bar();
        };
        // deals virtual rows to the warps, longest first, each to the least loaded warp
        auto deal = [&](std::vector<int> vs) {
            std::stable_sort(vs.begin(), vs.end(), [&](int x, int y) { return vrows[(size_t) x].ent.size() > vrows[(size_t) y].ent.size(); });
            int64_t load[kFrontMaxWarps] = {0};
            for (auto &v : of_warp) v.clear();
            for (int v : vs) {
                int best = 0;
                for (int wi = 1; wi < NW; ++wi)
                    if (load[wi] < load[best]) best = wi;
                of_warp[(size_t) best].push_back(v);
                load[best] += vcost(v);
            }
            emit_phase();
        };
        auto subtree_phase = [&]() {
            phase_kind = 0;
            for (auto &v : of_warp) v.clear();
            for (int q = 0; q < n_int; ++q) {
                const int r = upper ? n_int - 1 - q : q;
                if (group[(size_t) r] >= 0) of_warp[(size_t) warp_of_group[(size_t) group[(size_t) r]]].push_back(v_sub[(size_t) r]);
            }
            emit_phase();
        };
        auto upper_phases = [&]() {
            if (n_lv == 0) return;
            std::vector<int> as;
            for (int r = 0; r < n_int; ++r)
                if (v_a[(size_t) r] >= 0) as.push_back(v_a[(size_t) r]);
            phase_kind = 1;
            deal(as);
            phase_kind = 2;
            // (B) as one dense step: the upper rows' coupling among themselves is a small triangular system T x = v
            // (T = the entries between upper rows and the pivots, in processing order); x = inv(T) v is a dense
            // product every warp does for its rows at once (FR_LTOP), instead of a chain of short dependent rows.
            std::vector<int> tops;
            for (int q = 0; q < n_int; ++q) {
                const int r = upper ? n_int - 1 - q : q;
                if (level[(size_t) r] >= 0) tops.push_back(r);
            }
            const int T = (int) tops.size();
            if (T >= 2 && T <= 4 * NW) {
                std::vector<int> pos((size_t) n_int, -1);
                for (int pq = 0; pq < T; ++pq) pos[(size_t) tops[(size_t) pq]] = pq;
                std::vector<double> M((size_t) T * T, 0.0), inv((size_t) T * T, 0.0);
                for (int pq = 0; pq < T; ++pq) {
                    const int r = tops[(size_t) pq];
                    M[(size_t) pq * T + pq] = diag[(size_t) r];
                    for (const Entry &e : vrows[(size_t) v_b[(size_t) r]].ent) {
                        const int qq = pos[(size_t) e.idx];
                        if (qq < 0 || qq >= pq) throw std::logic_error("front schedule: upper rows of a local task are not triangular");
                        M[(size_t) pq * T + qq] += e.v;
                    }
                }
                for (int c = 0; c < T; ++c)
                    for (int i = c; i < T; ++i) {
                        double acc = i == c ? 1.0 : 0.0;
                        for (int q2 = c; q2 < i; ++q2) acc -= M[(size_t) i * T + q2] * inv[(size_t) q2 * T + c];
                        inv[(size_t) i * T + c] = acc / M[(size_t) i * T + i];
                    }
                // payload: T row indices (u8, padded to 16 bytes); then per warp T x 4 coefficients: [q][i] = inv[w + i n_warps][q]
                const size_t idx_bytes = round16((size_t) T);
                const size_t off = w.record(FR_LTOP, (uint32_t) T, (uint32_t) sj, 0, idx_bytes + (size_t) NW * T * 4 * (size_t) es);
                uint8_t idx[256];
                for (int pq = 0; pq < T; ++pq) idx[pq] = (uint8_t) tops[(size_t) pq];
                w.put_bytes(off, idx, (size_t) T);
                for (int wi = 0; wi < NW; ++wi)
                    for (int qq = 0; qq < T; ++qq)
                        for (int i = 0; i < 4; ++i) {
                            const int pq = wi + i * NW;
                            w.put_value(off + idx_bytes, ((size_t) wi * T + qq) * 4 + i, pq < T ? inv[(size_t) pq * T + qq] : 0.0);
                        }
                img.n_stored_values += (int64_t) NW * T * 4;
                img.local_phase_groups[2] += (T + 3) / 4;
                img.local_phase_rows[2] += 1;
                emit_bar();
                return;
            }
            if (upper_serial) {
                for (auto &v : of_warp) v.clear();
                for (int q = 0; q < n_int; ++q) {
                    const int r = upper ? n_int - 1 - q : q;
                    if (level[(size_t) r] >= 0) of_warp[0].push_back(v_b[(size_t) r]);
                }
                emit_phase();
                return;
            }
            for (int lv = 0; lv < n_lv; ++lv) {
                std::vector<int> vs;
                for (int r = 0; r < n_int; ++r)
                    if (level[(size_t) r] == lv) vs.push_back(v_b[(size_t) r]);
                deal(vs);
            }
        };
        if (!upper) {
            subtree_phase();
            upper_phases();
            phase_kind = 3;
            deal(v_contrib);
        } else {
            upper_phases();
            subtree_phase();
        }
        if (vrows.empty()) emit_bar();
        {
            const size_t off = w.record(FR_LSTORE, (uint32_t) n_int | (uint32_t) n_out << 16, (uint32_t) sj, (uint32_t) loc.c_lo,
                                        16 + (upper ? 4 * pad4((size_t) n_int) : 0));
            if (w0 >= (int64_t) 1 << 31) throw std::logic_error("front schedule: update-row buffer too large");
            w.put_int(off, 0, (int32_t) w0);
            if (upper)
                for (int i = 0; i < n_int; ++i) w.put_int(off, 4 + (size_t) i, L.perm[(size_t) (loc.c_lo + i)]);
        }
        emit_rel(sj);
        local_task[(size_t) li] = cur;
        for (int q = loc.sn_lo; q <= loc.sn_hi; ++q) done_of[(size_t) q] = {cur};
        close_task();
    }

    // task owning (publishing) a dense row in this sweep
    int owner_task(int row) const {
        const int s = S.sn_of[(size_t) row];
        if (S.local_of[(size_t) s] >= 0) return local_task[(size_t) S.local_of[(size_t) s]];
        return panel_task[(size_t) s][(size_t) ((row - S.c0[s]) / P)];
    }
    int64_t alloc_w(int rows) {
        const int64_t at = img.n_w_rows;
        img.n_w_rows += rows;
        if (img.n_w_rows >= (int64_t) 1 << 31) throw std::logic_error("front schedule: update-row buffer too large");
        return at;
    }

    // -- dense supernodes -------------------------------------------------------------------------
    // Source chunks of panel [p0, p1) of supernode s, in the order they are consumed.
    void panel_sources(int s, int p0, int p1, std::vector<Chunk> &out) const {
        out.clear();
        const int c0 = S.c0[s], c1 = S.c0[s + 1];
        if (!upper) {
            for (int a = c0; a < p0; a += kFrontLoad)  // earlier columns of the supernode, ascending
                out.push_back(Chunk{false, a, std::min(kFrontLoad, p0 - a), panel_task[(size_t) s][(size_t) ((std::min(a + kFrontLoad, p0) - 1 - c0) / P)], 0});
        } else {
            const int m = below(s);
            for (int q = 0; q < m; q += kFrontLoad)  // below-rows first (their owners finished long ago) ...
                out.push_back(Chunk{true, q, std::min(kFrontLoad, m - q), q == 0 ? owner_task(below_rows(s)[0]) : -1, 3});
            for (int pa = ((c1 - c0 - 1) / P) * P + c0; pa >= p1; pa -= P)  // ... then the later panels, nearest last
                for (int a = pa; a < std::min(pa + P, c1); a += kFrontLoad)
                    out.push_back(Chunk{false, a, std::min(kFrontLoad, std::min(pa + P, c1) - a), panel_task[(size_t) s][(size_t) ((pa - c0) / P)], 2});
        }
    }

    void emit_dense(int s) {
        const int c0 = S.c0[s], wd = width(s), m = below(s);
        const int np = (wd + P - 1) / P;
        panel_task[(size_t) s].assign((size_t) np, -1);
        std::vector<int> child_waits;
        if (!upper)
            for (int c : children[(size_t) s]) child_waits.insert(child_waits.end(), done_of[(size_t) c].begin(), done_of[(size_t) c].end());
        std::vector<Chunk> src, part;
        std::vector<std::pair<int, int32_t>> entries;
        std::vector<int> done;
        // Forward, wide supernodes: the right-hand side of every panel is ASSEMBLED AHEAD of the panel chain by a task of
        // its own (v = b - update rows, parked in the panel's rows of the work vector): the update rows of all the
        // descendants are then subtracted off the critical path (measured on cfg4: the assembly of a top-separator
        // panel, thousands of row loads, took 150 us of every 170 us chain step).
        std::vector<int> assembly_task;
        if (!upper && np > 1) {
            for (int pi = 0; pi < np; ++pi) {
                const int p0 = c0 + pi * P, p1 = std::min(p0 + P, c0 + wd), pw = p1 - p0;
                entries.clear();
                for (int r = 0; r < pw; ++r) {
                    std::vector<int32_t> &lst = contrib[(size_t) dense_row[(size_t) (p0 + r)]];
                    for (int32_t wr : lst) entries.push_back({r, wr});
                    std::vector<int32_t>().swap(lst);
                }
                open_task(s);
                img.n_dense_tasks += 1;
                const int sj = emit_loadx(0, 0, 0, 2, &L.perm[(size_t) p0], pw, 0, child_waits);
                emit_asm(sj, 0, entries);
                emit_bar();
                const size_t off = w.record(FR_LSTORE, (uint32_t) pw, (uint32_t) sj, (uint32_t) p0, 16);
                w.put_int(off, 0, 0);
                emit_rel(sj);
                assembly_task.push_back(cur);
                close_task();
            }
        }
        for (int step = 0; step < np; ++step) {
            const int pi = upper ? np - 1 - step : step;
            const int p0 = c0 + pi * P, p1 = std::min(p0 + P, c0 + wd), pw = p1 - p0;
            const int nt = (pw + TR - 1) / TR;
            panel_sources(s, p0, p1, src);
            // partial-product tasks for all but the trailing own_chunks chunks
            const int n_own = (int) std::min<size_t>(src.size(), (size_t) opt.own_chunks);
            const int n_far = (int) src.size() - n_own;
            std::vector<int> part_tasks;
            entries.clear();
            for (int a = 0; a < n_far; a += opt.split_chunks) {
                const int b = std::min(n_far, a + opt.split_chunks);
                part.assign(src.begin() + a, src.begin() + b);
                // each chunk waits for its own producer (a -1 relied on an earlier chunk of the full list)
                if (upper && part[0].wait < 0 && part[0].gather) part[0].wait = owner_task(below_rows(s)[0]);
                open_task(s);
                img.n_dense_tasks += 1;
                emit_accumulate(s, p0, pw, part, true);
                const int64_t w0 = alloc_w(pw);
                emit_epi(nt, FE_W, pw, 0, 0, (int) w0, nullptr);
                for (int r = 0; r < pw; ++r) entries.push_back({r, (int32_t) (w0 + r)});
                part_tasks.push_back(cur);
                close_task();
            }
            // the panel task itself
            open_task(s);
            img.n_dense_tasks += 1;
            part.assign(src.begin() + n_far, src.end());
            if (upper && !part.empty() && part[0].wait < 0 && part[0].gather) part[0].wait = owner_task(below_rows(s)[0]);
            emit_accumulate(s, p0, pw, part, true);
            std::vector<int> waits(part_tasks);
            int sjb;
            if (!upper && np > 1) {
                waits.push_back(assembly_task[(size_t) pi]);
                sjb = emit_loadx(1, p0, pw, 0, nullptr, 0, 0, waits);
            } else if (!upper) {
                waits.insert(waits.end(), child_waits.begin(), child_waits.end());
                sjb = emit_loadx(0, 0, 0, 2, &L.perm[(size_t) p0], pw, 0, waits);
                for (int r = 0; r < pw; ++r) {
                    std::vector<int32_t> &lst = contrib[(size_t) dense_row[(size_t) (p0 + r)]];
                    for (int32_t wr : lst) entries.push_back({r, wr});
                    std::vector<int32_t>().swap(lst);
                }
            } else {
                // the panel's right-hand side y (written by the forward sweep; nothing to wait for beyond the sources)
                if (m > 0 && part.empty() && part_tasks.empty()) waits.push_back(owner_task(below_rows(s)[0]));
                sjb = emit_loadx(1, p0, pw, 0, nullptr, 0, 0, waits);
            }
            if (!part.empty()) {
                emit_epi(nt, FE_RING_SUB, pw, sjb, 0, 0, nullptr);
                if (!entries.empty()) emit_bar();
            }
            emit_asm(sjb, 0, entries);
            emit_bar();
            emit_inverse(p0, pw, sjb, 0);
            // forward, single panel with a short below-product: keep it in this task (y stays in shared memory)
            const int n_groups = (m + NW * TR - 1) / (NW * TR);
            const int n_chunks = (wd + kFrontLoad - 1) / kFrontLoad;  // staged chunks over the supernode's columns
            const bool fuse = !upper && np == 1 && m > 0 && n_groups * ((wd + kFrontChunk - 1) / kFrontChunk) <= opt.fuse_steps;
            if (upper) {
                emit_epi(nt, FE_Y | FE_DST, pw, sjb, 0, p0, &L.perm[(size_t) p0]);
            } else if (!fuse) {
                emit_epi(nt, FE_Y, pw, sjb, 0, p0, nullptr);
            } else {
                emit_bar();
                emit_epi(nt, FE_Y | FE_RING_SET, pw, sjb, 0, p0, nullptr);
                emit_bar();
                const int64_t w0 = alloc_w(m);
                for (int g = 0; g < n_groups; ++g) {
                    const int q0 = g * NW * TR, rows = std::min(NW * TR, m - q0);
                    bool first = true;
                    for (int a = 0; a < wd; a += kFrontChunk) {
                        Chunk c{false, c0 + a, std::min(kFrontChunk, wd - a), -1, 1};
                        const size_t off = emit_mma(0, (rows + TR - 1) / TR, c.n, first, sjb, a);
                        fill_block(off, s, 1, q0, rows, c);
                        first = false;
                    }
                    emit_epi((rows + TR - 1) / TR, FE_W, rows, 0, 0, (int) (w0 + q0), nullptr);
                }
                for (int q = 0; q < m; ++q) contrib[(size_t) dense_row[(size_t) below_rows(s)[q]]].push_back((int32_t) (w0 + q));
            }
            emit_rel(sjb);
            panel_task[(size_t) s][(size_t) pi] = cur;
            if (upper || step == np - 1) done.push_back(cur);
            const bool fused = fuse;
            close_task();

            // forward: the below-product U = L[R, C] y as row groups x column ranges, each its own task
            if (!upper && step == np - 1 && m > 0 && !fused) {
                const int n_ranges = (n_chunks + opt.split_chunks - 1) / opt.split_chunks;
                for (int rg = 0; rg < n_ranges; ++rg) {
                    const int a0 = rg * opt.split_chunks * kFrontLoad, a1 = std::min(wd, a0 + opt.split_chunks * kFrontLoad);
                    const int64_t w0 = alloc_w(m);
                    for (int g = 0; g < n_groups; ++g) {
                        const int q0 = g * NW * TR, rows = std::min(NW * TR, m - q0);
                        open_task(s);
                        img.n_dense_tasks += 1;
                        part.clear();
                        for (int a = a0; a < a1; a += kFrontLoad)
                            part.push_back(Chunk{false, c0 + a, std::min(kFrontLoad, a1 - a), panel_task[(size_t) s][(size_t) ((std::min(a + kFrontLoad, a1) - 1) / P)], 1});
                        emit_accumulate(s, q0, rows, part, true);
                        emit_epi((rows + TR - 1) / TR, FE_W, rows, 0, 0, (int) (w0 + q0), nullptr);
                        done.push_back(cur);
                        close_task();
                    }
                    for (int q = 0; q < m; ++q) contrib[(size_t) dense_row[(size_t) below_rows(s)[q]]].push_back((int32_t) (w0 + q));
                }
            }
        }
        done_of[(size_t) s] = done;
    }

    void run() {
        const int ns = S.ns;
        done_of.assign((size_t) ns, {});
        panel_task.assign((size_t) ns, {});
        local_task.assign(S.locals.size(), -1);
        children.assign((size_t) ns, {});
        for (int s = 0; s < ns; ++s)
            if (S.parent[s] >= 0) children[(size_t) S.parent[s]].push_back(s);
        dense_row.assign((size_t) S.n, -1);
        int nd = 0;
        for (int s = 0; s < ns; ++s)
            if (S.local_of[s] < 0)
                for (int c = S.c0[s]; c < S.c0[s + 1]; ++c) dense_row[(size_t) c] = nd++;
        if (!upper) contrib.assign((size_t) nd, {});
        if (!upper) {
            int s = 0;
            while (s < ns) {
                if (S.local_of[s] >= 0) {
                    emit_local(S.local_of[s]);
                    s = S.locals[(size_t) S.local_of[s]].sn_hi + 1;
                } else {
                    emit_dense(s);
                    ++s;
                }
            }
        } else {
            int s = ns - 1;
            while (s >= 0) {
                if (S.local_of[s] >= 0) {
                    emit_local(S.local_of[s]);
                    s = S.locals[(size_t) S.local_of[s]].sn_lo - 1;
                } else {
                    emit_dense(s);
                    --s;
                }
            }
        }
    }
};

}  // namespace

void build_front_image(const HostFactor &L, const FrontStructure &S, int elem, bool upper, const FrontOptions &opt,
                       FrontImage *out) {
    FrontImage &img = *out;
    img = FrontImage();
    img.upper = upper;
    img.elem = elem;
    img.n = L.n;
    img.tile_rows = front_tile_rows(elem);
    img.n_warps = opt.warps;
    img.page_bytes = front_page_bytes(elem, opt.warps);
    img.slab.reserve((size_t) ((double) L.nnz() * elem * 1.15) + (size_t) S.ns * 64 + 65536);
    FrontBuilder b(L, S, opt, img);
    b.run();

    // Ticket order: subtree block by subtree block in postorder (forward; mirrored backward, the top block first),
    // by dependency level inside a block, larger slabs first inside a level.  Every dependency of a task lies in its
    // own block at a lower level, or (backward) in the top block: a topological order.
    const int nt = (int) img.tasks.size();
    uint32_t max_level = 0;
    for (const TaskDesc &t : img.tasks) max_level = std::max(max_level, t.level);
    img.n_levels = nt ? (int) max_level + 1 : 0;
    img.order.resize((size_t) nt);
    std::iota(img.order.begin(), img.order.end(), 0);
    const int top = S.n_blocks - 1;
    auto block_key = [&](int t) {
        const int blk = b.block_of_task[(size_t) t];
        if (!upper) return blk;
        return blk == top ? -1 : top - blk;
    };
    std::stable_sort(img.order.begin(), img.order.end(), [&](int a, int c) {
        const int ka = block_key(a), kc = block_key(c);
        if (ka != kc) return ka < kc;
        const TaskDesc &x = img.tasks[(size_t) a], &y = img.tasks[(size_t) c];
        if (x.level != y.level) return x.level < y.level;
        return x.bytes > y.bytes;
    });
}

}  // namespace chb
