// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
//
// Sequential CPU interpreter of the task/record stream that cholespy_b200's host code builds for the
// sm_100a sweep kernels (cholespy_b200/csrc/panel_schedule.h).  It executes the tasks one at a time
// in TICKET ORDER with the arithmetic the device kernels use (fused multiply-subtract, IEEE
// division, same accumulation order), and fails if a task consumes something a later ticket
// produces.  The CPU tests use it to check the schedule builder against the row-by-row oracle
// (cholespy_oracle.c, which follows the reference kernels/kernels.cu:81-120,149-190) without a GPU.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int W = 16;

struct TaskDesc {
    uint64_t offset;
    uint32_t bytes;
    uint32_t level;
};

inline float fnma(float a, float b, float c) { return __builtin_fmaf(-a, b, c); }
inline double fnma(double a, double b, double c) { return __builtin_fma(-a, b, c); }
inline int span16(uint32_t m) { return m ? 32 - __builtin_clz(m) : 0; }

enum { NOP = 0, BEGIN_RHS, BEGIN_ZERO, BEGIN_SLOT, ADD_SLOT, BLOCK_D4, BLOCK_D1, BLOCK_TM, END_FINAL, END_SLOT, END_DENSE,
       LOCAL_LOAD, LOCAL_STORE, END_STORE, LOCAL_ROWS, ADD_ROWS, LOCAL_FLAG = 0x40, SRC_R_FLAG = 0x80 };

// returns 0, or a negative error code
template <typename F>
int run(const uint8_t *slab, const TaskDesc *tasks, const int32_t *order, int n_tasks, int page, int upper, int k, F *tmp,
        const F *b, F *x, F *scratch) {
    std::vector<uint8_t> done((size_t) n_tasks, 0);
    std::vector<F> acc((size_t) W * k);
    std::vector<F> cbuf;  // contribution buffer of the forward sweep (grown on demand)
    std::vector<F> rbuf;  // residual vector of the inverse super-blocks (grown on demand)
    std::vector<F> loc;  // rows of the local task being executed: [n_int unknowns | n_ext external rows] x k
    int loc_rows = 0;
    auto waited = [&](uint32_t c) { return c == 0 || done[c - 1]; };
    for (int ticket = 0; ticket < n_tasks; ++ticket) {
        const int task = order[ticket];
        const uint8_t *base = slab + tasks[task].offset;
        const uint32_t total = tasks[task].bytes;
        if (total % 16) return -2;
        int row0 = -1, w = 0;
        for (uint32_t pg = 0; pg < total; pg += (uint32_t) page) {
            const uint32_t plen = total - pg < (uint32_t) page ? total - pg : (uint32_t) page;
            uint32_t p = 0;
            while (p < plen) {
                uint32_t h[4];
                std::memcpy(h, base + pg + p, 16);
                const uint32_t type = h[0] & 0x3f, len = (h[0] >> 8) * 16;
                const bool local = (h[0] & LOCAL_FLAG) != 0;
                const bool src_r = (h[0] & SRC_R_FLAG) != 0;
                if (len < 16 || p + len > plen) return -3;  // record straddles a page
                const uint8_t *pay = base + pg + p + 16;
                const uint8_t *src_idx = pay;  // local blocks: 16 local source-row indices first
                if (local && type >= BLOCK_D4 && type <= BLOCK_TM) pay += 16;
                const F *v = (const F *) pay;
                // row `s` (sweep order rank r) of a block's source panel
                auto source_row = [&](int src0, int s, int r) -> const F * {
                    if (src_r) return (size_t) (src0 + s + 1) * k <= rbuf.size() ? rbuf.data() + (size_t) (src0 + s) * k : nullptr;
                    if (!local) return tmp + (size_t) (src0 + s) * k;
                    if ((int) src_idx[r] >= loc_rows) return nullptr;
                    return loc.data() + (size_t) src_idx[r] * k;
                };
                switch (type) {
                case NOP: break;
                case LOCAL_LOAD: {
                    const int lo = (int) h[1], n_int = (int) (h[2] & 0xffff), n_ext = (int) (h[2] >> 16), n_wait = (int) h[3];
                    const int32_t *list = (const int32_t *) pay;
                    loc_rows = n_int + n_ext;
                    loc.assign((size_t) loc_rows * k, F(0));
                    if (!upper) {
                        if (n_wait) return -7;  // n_ext rows (zero here) receive the task's contributions
                        for (int i = 0; i < n_int; ++i)
                            for (int c = 0; c < k; ++c) loc[(size_t) i * k + c] = b[(size_t) list[i] * k + c];
                    } else {
                        const int32_t *waits = list + ((n_ext + 3) & ~3);
                        for (int e = 0; e < n_wait; ++e)
                            if (!waited((uint32_t) waits[e])) return -4;
                        for (int i = 0; i < n_int; ++i)
                            for (int c = 0; c < k; ++c) loc[(size_t) i * k + c] = tmp[(size_t) (lo + i) * k + c];
                        for (int e = 0; e < n_ext; ++e)
                            for (int c = 0; c < k; ++c) loc[(size_t) (n_int + e) * k + c] = tmp[(size_t) list[e] * k + c];
                    }
                } break;
                case LOCAL_ROWS: {
                    const int n_rows = (int) h[1];
                    const bool reciprocal = h[2] == 1;
                    const uint8_t *q = pay;
                    for (int i = 0; i < n_rows; ++i) {
                        uint32_t meta;
                        std::memcpy(&meta, q, 4);
                        const int row = (int) (meta & 0xff), n4 = (int) (meta >> 8 & 0xff);
                        const bool from_zero = (meta >> 16 & 1) != 0;
                        F piv;
                        std::memcpy(&piv, q + 8, sizeof(F));
                        const uint8_t *idx = q + 16;
                        const F *val = (const F *) (q + 16 + ((4 * n4 + 15) & ~15));
                        if (row >= loc_rows) return -7;
                        for (int c = 0; c < k; ++c) {
                            F acc = from_zero ? F(0) : loc[(size_t) row * k + c];
                            for (int e = 0; e < 4 * n4; ++e) {
                                if ((int) idx[e] >= loc_rows) return -7;
                                acc = fnma(val[e], loc[(size_t) idx[e] * k + c], acc);
                            }
                            loc[(size_t) row * k + c] = reciprocal ? acc * piv : acc / piv;
                        }
                        q = (const uint8_t *) (val + 4 * n4);
                    }
                } break;
                case ADD_ROWS: {
                    if (!waited(h[3])) return -4;
                    const size_t first = (size_t) h[1];
                    const int nt = (int) (h[2] & 0xffff);
                    const uint32_t tmask = h[2] >> 16;
                    if ((first + nt) * k > cbuf.size()) return -7;
                    int r = 0;
                    for (int t = 0; t < W; ++t) {
                        if (!(tmask >> t & 1)) continue;
                        for (int c = 0; c < k; ++c) acc[(size_t) t * k + c] += cbuf[(first + r) * k + c];
                        ++r;
                    }
                    if (r != nt) return -5;
                } break;
                case LOCAL_STORE: {
                    const int lo = (int) h[1], n_int = (int) (h[2] & 0xffff), n_out = (int) (h[2] >> 16);
                    if (n_out) {
                        const size_t first = (size_t) h[3];
                        if (cbuf.size() < (first + n_out) * k) cbuf.resize((first + n_out) * k, F(0));
                        for (int e = 0; e < n_out; ++e)
                            for (int c = 0; c < k; ++c) cbuf[(first + e) * k + c] = loc[(size_t) (n_int + e) * k + c];
                    }
                    const int32_t *perm = (const int32_t *) pay;
                    for (int i = 0; i < n_int; ++i)
                        for (int c = 0; c < k; ++c) {
                            tmp[(size_t) (lo + i) * k + c] = loc[(size_t) i * k + c];
                            if (upper) x[(size_t) perm[i] * k + c] = loc[(size_t) i * k + c];
                        }
                    loc_rows = 0;
                } break;
                case BEGIN_RHS: {
                    row0 = (int) h[1], w = (int) h[2];
                    std::fill(acc.begin(), acc.end(), F(0));
                    if (local) {
                        if (row0 + w > loc_rows) return -7;
                        for (int i = 0; i < w; ++i)
                            for (int c = 0; c < k; ++c) acc[(size_t) i * k + c] = loc[(size_t) (row0 + i) * k + c];
                        break;
                    }
                    const int32_t *perm = (const int32_t *) pay;
                    for (int i = 0; i < w; ++i)
                        for (int c = 0; c < k; ++c)
                            acc[(size_t) i * k + c] = upper ? tmp[(size_t) (row0 + i) * k + c] : b[(size_t) perm[i] * k + c];
                } break;
                case BEGIN_ZERO:
                    row0 = (int) h[1], w = (int) h[2];
                    std::fill(acc.begin(), acc.end(), F(0));
                    break;
                case BEGIN_SLOT: {
                    row0 = (int) h[1], w = (int) (h[2] & 0xff);
                    if (!waited(h[3])) return -4;
                    const F *s = scratch + (size_t) (h[2] >> 8) * W * k;
                    std::copy(s, s + (size_t) W * k, acc.begin());
                } break;
                case ADD_SLOT: {
                    if (!waited(h[3])) return -4;
                    const F *s = scratch + (size_t) (h[2] >> 8) * W * k;
                    for (size_t i = 0; i < (size_t) W * k; ++i) acc[i] += s[i];
                } break;
                case BLOCK_D4: {
                    if (!waited(h[3])) return -4;
                    const int src0 = (int) h[1];
                    const uint32_t smask = h[2] & 0xffff, tmask = h[2] >> 16;
                    int group_rank[4], ng = 0;
                    for (int g = 0; g < 4; ++g) group_rank[g] = (tmask >> (4 * g) & 0xf) ? ng++ : -1;
                    const int stride = 4 * ng;
                    int r = 0;
                    for (int bit = 0; bit < W; ++bit) {
                        const int s = upper ? W - 1 - bit : bit;
                        if (!(smask >> s & 1)) continue;
                        const F *srow = source_row(src0, s, r);
                        if (!srow) return -7;
                        for (int t = 0; t < W; ++t) {
                            if (!(tmask >> t & 1)) continue;
                            const F l = v[(size_t) r * stride + 4 * group_rank[t >> 2] + (t & 3)];
                            for (int c = 0; c < k; ++c) acc[(size_t) t * k + c] = fnma(l, srow[c], acc[(size_t) t * k + c]);
                        }
                        ++r;
                    }
                } break;
                case END_FINAL: {
                    if ((int) h[1] != row0 || (int) h[2] != w) return -5;
                    const int32_t *perm = (const int32_t *) pay;
                    if (upper && !local) v = (const F *) (pay + 4 * W);
                    size_t idx = 0;
                    for (int step = 0; step < w; ++step) {
                        const int s = upper ? w - 1 - step : step;
                        const F d = v[idx++];
                        for (int c = 0; c < k; ++c) acc[(size_t) s * k + c] = acc[(size_t) s * k + c] / d;
                        const int t0 = upper ? 0 : s + 1, t1 = upper ? s : w;
                        for (int t = t0; t < t1; ++t) {
                            const F l = v[idx++];
                            for (int c = 0; c < k; ++c) acc[(size_t) t * k + c] = fnma(l, acc[(size_t) s * k + c], acc[(size_t) t * k + c]);
                        }
                    }
                    for (int i = 0; i < w; ++i)
                        for (int c = 0; c < k; ++c) {
                            if (local) {
                                loc[(size_t) (row0 + i) * k + c] = acc[(size_t) i * k + c];
                                continue;
                            }
                            tmp[(size_t) (row0 + i) * k + c] = acc[(size_t) i * k + c];
                            if (upper) x[(size_t) perm[i] * k + c] = acc[(size_t) i * k + c];
                        }
                } break;
                case END_DENSE: {
                    if ((int) h[1] != row0 || (int) h[2] != w) return -5;
                    const int32_t *perm = (const int32_t *) pay;
                    if (upper && !local) v = (const F *) (pay + 4 * W);
                    const int wp = w == 1 ? 1 : (w + 3) & ~3;
                    for (int step = 0; step < wp; ++step) {
                        const int s = upper ? wp - 1 - step : step;
                        const F *m = v + (size_t) s * wp;
                        for (int c = 0; c < k; ++c) acc[(size_t) s * k + c] = acc[(size_t) s * k + c] * m[s];
                        const int t0 = upper ? 0 : s + 1, t1 = upper ? s : wp;
                        for (int t = t0; t < t1; ++t)
                            for (int c = 0; c < k; ++c) acc[(size_t) t * k + c] = fnma(m[t], acc[(size_t) s * k + c], acc[(size_t) t * k + c]);
                    }
                    for (int i = 0; i < w; ++i)
                        for (int c = 0; c < k; ++c) {
                            if (local) {
                                loc[(size_t) (row0 + i) * k + c] = acc[(size_t) i * k + c];
                                continue;
                            }
                            tmp[(size_t) (row0 + i) * k + c] = acc[(size_t) i * k + c];
                            if (upper) x[(size_t) perm[i] * k + c] = acc[(size_t) i * k + c];
                        }
                } break;
                case END_STORE: {
                    if ((int) h[1] != row0 || (int) h[2] != w) return -5;
                    const int32_t *perm = (const int32_t *) pay;
                    if (h[3] == 1) {
                        if (rbuf.size() < (size_t) (row0 + w) * k) rbuf.resize((size_t) (row0 + w) * k, F(0));
                        for (int i = 0; i < w; ++i)
                            for (int c = 0; c < k; ++c) rbuf[(size_t) (row0 + i) * k + c] = acc[(size_t) i * k + c];
                    } else {
                        for (int i = 0; i < w; ++i)
                            for (int c = 0; c < k; ++c) {
                                tmp[(size_t) (row0 + i) * k + c] = acc[(size_t) i * k + c];
                                if (upper) x[(size_t) perm[i] * k + c] = acc[(size_t) i * k + c];
                            }
                    }
                } break;
                case END_SLOT: {
                    F *s = scratch + (size_t) (h[2] >> 8) * W * k;
                    std::copy(acc.begin(), acc.end(), s);
                } break;
                default: return -6;
                }
                p += len;
            }
        }
        done[task] = 1;
    }
    return 0;
}

}  // namespace

extern "C" int chb_emu_sweep_f32(const void *slab, const void *tasks, const int32_t *order, int n_tasks, int page, int upper, int k,
                                 float *tmp, const float *b, float *x, float *scratch) {
    return run<float>((const uint8_t *) slab, (const TaskDesc *) tasks, order, n_tasks, page, upper, k, tmp, b, x, scratch);
}
extern "C" int chb_emu_sweep_f64(const void *slab, const void *tasks, const int32_t *order, int n_tasks, int page, int upper, int k,
                                 double *tmp, const double *b, double *x, double *scratch) {
    return run<double>((const uint8_t *) slab, (const TaskDesc *) tasks, order, n_tasks, page, upper, k, tmp, b, x, scratch);
}
