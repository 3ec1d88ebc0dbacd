// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
//
// Sequential CPU interpreter of the "front" task/record stream that cholespy_b200's host code builds for the
// sm_100a front sweep kernels (cholespy_b200/csrc/front_schedule.h).  It executes the tasks one at a time in
// TICKET ORDER with the arithmetic of the device kernels (fused multiply-add in source order, reciprocal
// pivots), and fails if a task waits for, or reads a row written by, a task with a later ticket.  The CPU tests
// compare its result with the row-by-row oracle (cholespy_oracle.c, which follows the reference
// kernels/kernels.cu:81-120,149-190) without a GPU.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct TaskDesc {
    uint64_t offset;
    uint32_t bytes;
    uint32_t level;
};

enum { FR_NOP = 0, FR_LOADX, FR_MMA, FR_EPI, FR_ASM, FR_BAR, FR_REL, FR_LROWS, FR_LSTORE, FR_DIR, FR_LTOP };
enum { FE_RING_SUB = 1, FE_RING_SET = 2, FE_Y = 4, FE_W = 8, FE_DST = 16 };
constexpr int kMaxWarps = 8;

inline float fmadd(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline double fmadd(double a, double b, double c) { return __builtin_fma(a, b, c); }

template <typename F>
int run(const uint8_t *slab, const TaskDesc *tasks, const int32_t *order, int n_tasks, int page, int TR, int kWarps, int upper, int k, int n,
        int64_t n_w_rows, F *Y, F *W, const F *b, F *x) {
    std::vector<uint8_t> done((size_t) n_tasks, 0);
    std::vector<uint8_t> w_written((size_t) n_w_rows, 0), y_written((size_t) n, upper ? 1 : 0);
    if (kWarps < 1 || kWarps > kMaxWarps) return -20;
    std::vector<F> acc((size_t) kWarps * TR * k);
    for (int ticket = 0; ticket < n_tasks; ++ticket) {
        const int task = order[ticket];
        if (task < 0 || task >= n_tasks || done[(size_t) task]) return -1;
        const uint8_t *base = slab + tasks[task].offset;
        const uint32_t total = tasks[task].bytes;
        if (total % 16) return -2;
        std::vector<std::vector<F>> staged;   // rows of each staged record
        std::vector<uint8_t> released;
        for (uint32_t pg = 0; pg < total; pg += (uint32_t) page) {
            const uint32_t plen = total - pg < (uint32_t) page ? total - pg : (uint32_t) page;
            uint32_t p = 0;
            while (p < plen) {
                uint32_t h[4];
                std::memcpy(h, base + pg + p, 16);
                const uint32_t type = h[0] & 0xff, len = (h[0] >> 8 & 0xffff) * 16;
                if ((p == 0) != (type == FR_DIR)) return -17;  // a directory record starts every page, and only there
                if (len < 16 || p + len > plen) return -3;  // record straddles a page
                const uint8_t *pay = base + pg + p + 16;
                const int32_t *ipay = (const int32_t *) pay;
                auto stage = [&](uint32_t sj) -> std::vector<F> * {
                    if (sj >= staged.size() || released[sj]) return nullptr;
                    return &staged[sj];
                };
                switch (type) {
                case FR_NOP: case FR_BAR: break;
                case FR_DIR: {
                    // the directory must list exactly the page's staged records
                    const uint32_t cnt = h[0] >> 24;
                    uint32_t found = 0, q = 16;
                    while (q < plen) {
                        uint32_t g0;
                        std::memcpy(&g0, base + pg + q, 4);
                        if ((g0 & 0xff) == FR_LOADX) {
                            if (found < 6 && cnt <= 6) {
                                uint16_t units;
                                std::memcpy(&units, base + pg + 4 + 2 * found, 2);
                                if ((uint32_t) units * 16 != q) return -18;
                            }
                            found += 1;
                        }
                        if ((g0 >> 8 & 0xffff) == 0) return -18;
                        q += (g0 >> 8 & 0xffff) * 16;
                    }
                    if (cnt != (found > 6 ? 7u : found)) return -18;
                    break;
                }
                case FR_LOADX: {
                    const int n_contig = (int) (h[1] & 1023), n_gather = (int) (h[1] >> 10 & 1023), n_extra = (int) (h[1] >> 20 & 1023);
                    const int row0 = (int) h[2];
                    const int cbuf = (int) (h[3] & 15), gkind = (int) (h[3] >> 4 & 15), n_wait = (int) (h[3] >> 8);
                    for (int i = 0; i < n_wait; ++i) {
                        if (ipay[i] < 0 || ipay[i] >= n_tasks || !done[(size_t) ipay[i]]) return -4;  // producer has a later ticket
                    }
                    const int32_t *gl = ipay + ((n_wait + 3) & ~3);
                    std::vector<F> rows((size_t) (n_contig + n_gather + n_extra) * k, F(0));
                    for (int i = 0; i < n_contig; ++i) {
                        const int64_t r = (int64_t) row0 + i;
                        const F *src;
                        if (cbuf == 1) {
                            if (r < 0 || r >= n || !y_written[(size_t) r]) return -5;
                            src = Y + (size_t) r * k;
                        } else if (cbuf == 2) {
                            if (r < 0 || r >= n_w_rows || !w_written[(size_t) r]) return -5;
                            src = W + (size_t) r * k;
                        } else return -6;
                        std::memcpy(&rows[(size_t) i * k], src, sizeof(F) * (size_t) k);
                    }
                    for (int i = 0; i < n_gather; ++i) {
                        const int r = gl[i];
                        if (r < 0 || r >= n) return -5;
                        const F *src;
                        if (gkind == 1) {
                            if (!y_written[(size_t) r]) return -5;
                            src = Y + (size_t) r * k;
                        } else if (gkind == 2) {
                            src = b + (size_t) r * k;
                        } else return -6;
                        std::memcpy(&rows[(size_t) (n_contig + i) * k], src, sizeof(F) * (size_t) k);
                    }
                    staged.push_back(std::move(rows));
                    released.push_back(0);
                    break;
                }
                case FR_MMA: {
                    const int tile0 = (int) (h[1] & 255), n_tiles = (int) (h[1] >> 8 & 255), n_src = (int) (h[1] >> 16 & 255);
                    const bool zero = (h[1] >> 24 & 1) != 0;
                    std::vector<F> *st = stage(h[2] & 0xffff);
                    const int src_row = (int) (h[2] >> 16);
                    if (!st || (size_t) (src_row + n_src) * k > st->size() || tile0 + n_tiles > kWarps) return -7;
                    const F *A = (const F *) pay;
                    for (int t = 0; t < n_tiles; ++t) {
                        F *a = &acc[(size_t) (tile0 + t) * TR * k];
                        if (zero) std::fill(a, a + (size_t) TR * k, F(0));
                        for (int s = 0; s < n_src; ++s) {
                            const F *xs = st->data() + (size_t) (src_row + s) * k;
                            const F *av = A + ((size_t) t * n_src + s) * TR;
                            for (int i = 0; i < TR; ++i)
                                for (int c = 0; c < k; ++c) a[(size_t) i * k + c] = fmadd(av[i], xs[c], a[(size_t) i * k + c]);
                        }
                    }
                    break;
                }
                case FR_EPI: {
                    const int tile0 = (int) (h[1] & 255), n_tiles = (int) (h[1] >> 8 & 255);
                    const uint32_t mode = h[1] >> 16;
                    const int n_rows = (int) (h[2] & 0xffff);
                    const int64_t g0 = (int64_t) (int32_t) h[3];
                    const int ring_row = ipay[0];
                    if (n_rows > n_tiles * TR) return -8;
                    std::vector<F> *st = nullptr;
                    if (mode & (FE_RING_SUB | FE_RING_SET)) {
                        st = stage(h[2] >> 16);
                        if (!st || (size_t) (ring_row + n_rows) * k > st->size()) return -8;
                    }
                    for (int r = 0; r < n_rows; ++r) {
                        const F *a = &acc[((size_t) tile0 * TR + r) * k];
                        for (int c = 0; c < k; ++c) {
                            const F v = a[c];
                            if (mode & FE_RING_SUB) (*st)[(size_t) (ring_row + r) * k + c] -= v;
                            if (mode & FE_RING_SET) (*st)[(size_t) (ring_row + r) * k + c] = v;
                            if (mode & FE_Y) Y[(size_t) (g0 + r) * k + c] = v;
                            if (mode & FE_W) W[(size_t) (g0 + r) * k + c] = v;
                            if (mode & FE_DST) x[(size_t) ipay[4 + r] * k + c] = v;
                        }
                        if (mode & FE_Y) {
                            if (g0 + r < 0 || g0 + r >= n) return -8;
                            y_written[(size_t) (g0 + r)] = 1;
                        }
                        if (mode & FE_W) {
                            if (g0 + r < 0 || g0 + r >= n_w_rows) return -8;
                            w_written[(size_t) (g0 + r)] = 1;
                        }
                    }
                    break;
                }
                case FR_ASM: {
                    std::vector<F> *st = stage(h[1] & 0xffff);
                    const int ring_row = (int) (h[1] >> 16);
                    const int n_entries = (int) h[2];
                    if (!st) return -9;
                    if (ipay[kWarps] != n_entries) return -9;
                    const int32_t *ent = ipay + 12;
                    for (int wi = 0; wi < kWarps; ++wi)
                        for (int e = ipay[wi]; e < ipay[wi + 1]; ++e) {
                            const int r = ent[2 * e];
                            const int64_t wr = ent[2 * e + 1];
                            if (r % kWarps != wi) return -9;   // a ring row must belong to one warp
                            if ((size_t) (ring_row + r + 1) * k > st->size() || wr < 0 || wr >= n_w_rows || !w_written[(size_t) wr]) return -10;
                            for (int c = 0; c < k; ++c) (*st)[(size_t) (ring_row + r) * k + c] -= W[(size_t) wr * k + c];
                        }
                    break;
                }
                case FR_REL: {
                    const uint32_t sj = h[1];
                    // staged records are released in staging order
                    if (sj >= staged.size() || released[sj]) return -11;
                    for (uint32_t q = 0; q < sj; ++q)
                        if (!released[q]) return -11;
                    released[sj] = 1;
                    std::vector<F>().swap(staged[sj]);
                    break;
                }
                case FR_LROWS: {
                    std::vector<F> *st = stage(h[2]);
                    if (!st) return -12;
                    const int rows_avail = (int) (st->size() / (size_t) k);
                    const size_t es = sizeof(F);
                    for (int wi = 0; wi < kWarps; ++wi) {
                        const uint8_t *q = pay + ipay[wi], *qend = pay + ipay[wi + 1];
                        if (ipay[wi] < 48 || ipay[wi + 1] < ipay[wi] || (uint32_t) ipay[wi + 1] > len - 16) return -12;
                        while (q < qend) {
                            uint32_t meta;
                            std::memcpy(&meta, q, 4);
                            F pivot;
                            std::memcpy(&pivot, q + es, sizeof(F));
                            const int n4 = (int) (meta >> 8 & 0xff), row = (int) (meta & 0xff);
                            const bool contribution = (meta >> 16 & 1) != 0;
                            const uint8_t *idx = q + 2 * es;
                            const F *val = (const F *) (q + ((2 * es + 4 * n4 + 15) & ~(size_t) 15));
                            if (row >= rows_avail) return -12;
                            for (int c = 0; c < k; ++c) {
                                F a = contribution ? F(0) : (*st)[(size_t) row * k + c];
                                for (int e = 0; e < 4 * n4; ++e) {
                                    if ((int) idx[e] >= rows_avail || (int) idx[e] == row) return -12;
                                    a = fmadd(-val[e], (*st)[(size_t) idx[e] * k + c], a);
                                }
                                (*st)[(size_t) row * k + c] = a * pivot;
                            }
                            q = (const uint8_t *) (val + 4 * n4);
                        }
                    }
                    break;
                }
                case FR_LTOP: {
                    const int T = (int) h[1];
                    std::vector<F> *st = stage(h[2]);
                    if (!st || T < 1 || T > 4 * kWarps) return -19;
                    const int rows_avail = (int) (st->size() / (size_t) k);
                    const uint8_t *idx = pay;
                    const F *coef = (const F *) (pay + ((T + 15) & ~15));
                    std::vector<F> out((size_t) T * k, F(0));
                    for (int pq = 0; pq < T; ++pq) {
                        const int wi = pq % kWarps, i = pq / kWarps;
                        if ((int) idx[pq] >= rows_avail) return -19;
                        for (int qq = 0; qq < T; ++qq) {
                            const F cf = coef[((size_t) wi * T + qq) * 4 + i];
                            for (int c = 0; c < k; ++c) out[(size_t) pq * k + c] = fmadd(cf, (*st)[(size_t) idx[qq] * k + c], out[(size_t) pq * k + c]);
                        }
                    }
                    for (int pq = 0; pq < T; ++pq) std::memcpy(&(*st)[(size_t) idx[pq] * k], &out[(size_t) pq * k], sizeof(F) * (size_t) k);
                    break;
                }
                case FR_LSTORE: {
                    const int n_int = (int) (h[1] & 0xffff), n_ext = (int) (h[1] >> 16);
                    std::vector<F> *st = stage(h[2]);
                    const int64_t y0 = (int64_t) (int32_t) h[3], w0 = ipay[0];
                    if (!st || (size_t) (n_int + n_ext) * k > st->size()) return -13;
                    for (int i = 0; i < n_int; ++i) {
                        if (y0 + i >= n) return -13;
                        for (int c = 0; c < k; ++c) {
                            const F v = (*st)[(size_t) i * k + c];
                            Y[(size_t) (y0 + i) * k + c] = v;
                            if (upper) x[(size_t) ipay[4 + i] * k + c] = v;
                        }
                        y_written[(size_t) (y0 + i)] = 1;
                    }
                    for (int e = 0; e < n_ext; ++e) {
                        if (w0 + e >= n_w_rows) return -13;
                        std::memcpy(W + (size_t) (w0 + e) * k, &(*st)[(size_t) (n_int + e) * k], sizeof(F) * (size_t) k);
                        w_written[(size_t) (w0 + e)] = 1;
                    }
                    break;
                }
                default: return -14;
                }
                p += len;
            }
        }
        for (size_t q = 0; q < released.size(); ++q)
            if (!released[q]) return -15;  // a task must release everything it staged
        done[(size_t) task] = 1;
    }
    if (!upper)
        for (int r = 0; r < n; ++r)
            if (!y_written[(size_t) r]) return -16;
    return 0;
}

}  // namespace

extern "C" {
int chb_front_emu_sweep_f32(const void *slab, const void *tasks, const int32_t *order, int n_tasks, int page, int tile_rows, int n_warps, int upper,
                            int k, int n, int64_t n_w_rows, float *Y, float *W, const float *b, float *x) {
    return run<float>((const uint8_t *) slab, (const TaskDesc *) tasks, order, n_tasks, page, tile_rows, n_warps, upper, k, n, n_w_rows, Y, W, b, x);
}
int chb_front_emu_sweep_f64(const void *slab, const void *tasks, const int32_t *order, int n_tasks, int page, int tile_rows, int n_warps, int upper,
                            int k, int n, int64_t n_w_rows, double *Y, double *W, const double *b, double *x) {
    return run<double>((const uint8_t *) slab, (const TaskDesc *) tasks, order, n_tasks, page, tile_rows, n_warps, upper, k, n, n_w_rows, Y, W, b, x);
}
}
