// Disk cache of factors and sweep images.  See disk_cache.h.
#include "disk_cache.h"

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace chb {

namespace {

constexpr uint32_t kMagic = 0x31424843u;  // "CHB1"
constexpr uint32_t kFormat = 4;           // bump whenever a serialised struct or a record layout changes

// 64-bit multiply-xorshift mixer over 8-byte words (two differently seeded lanes make the 128-bit key)
struct Hasher {
    uint64_t h;
    explicit Hasher(uint64_t seed) : h(seed) {}
    void word(uint64_t w) {
        h ^= w + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 33;
    }
    void bytes(const void *p, size_t n) {
        const unsigned char *c = (const unsigned char *) p;
        size_t i = 0;
        for (; i + 8 <= n; i += 8) {
            uint64_t w;
            std::memcpy(&w, c + i, 8);
            word(w);
        }
        uint64_t tail = 0;
        if (i < n) std::memcpy(&tail, c + i, n - i);
        word(tail ^ ((uint64_t) n << 56));
    }
};

CacheKey hash_arrays(int n, const std::vector<int> &ptr, const std::vector<int> &idx, const std::vector<double> &val, uint64_t salt) {
    CacheKey k;
    for (int lane = 0; lane < 2; ++lane) {
        Hasher h(lane ? 0x243f6a8885a308d3ull ^ salt : 0x13198a2e03707344ull + salt);
        h.word((uint64_t) n);
        h.bytes(ptr.data(), sizeof(int) * ptr.size());
        h.bytes(idx.data(), sizeof(int) * idx.size());
        h.bytes(val.data(), sizeof(double) * val.size());
        (lane ? k.b : k.a) = h.h;
    }
    return k;
}

struct File {
    FILE *f = nullptr;
    std::string final_path, tmp_path;
    bool ok = true;
    ~File() {
        if (f) fclose(f);
    }
    bool open_read(const std::string &path) {
        f = fopen(path.c_str(), "rb");
        return f != nullptr;
    }
    bool open_write(const std::string &path) {
        final_path = path;
        tmp_path = path + ".tmp" + std::to_string((long) getpid());
        f = fopen(tmp_path.c_str(), "wb");
        return f != nullptr;
    }
    void commit() {
        if (!f) return;
        const bool good = ok && fflush(f) == 0;
        fclose(f);
        f = nullptr;
        if (good) rename(tmp_path.c_str(), final_path.c_str());
        else remove(tmp_path.c_str());
    }
    template <typename T> void put(const T &v) { ok = ok && fwrite(&v, sizeof(T), 1, f) == 1; }
    template <typename T> void get(T &v) { ok = ok && fread(&v, sizeof(T), 1, f) == 1; }
    template <typename T> void put_vec(const std::vector<T> &v) {
        put((uint64_t) v.size());
        if (!v.empty()) ok = ok && fwrite(v.data(), sizeof(T), v.size(), f) == v.size();
    }
    template <typename T> void get_vec(std::vector<T> &v) {
        uint64_t n = 0;
        get(n);
        if (!ok || n > (1ull << 40) / sizeof(T)) {
            ok = false;
            return;
        }
        v.resize((size_t) n);
        if (n) ok = ok && fread(v.data(), sizeof(T), (size_t) n, f) == (size_t) n;
    }
    void put_header(uint32_t kind) {
        put(kMagic);
        put(kFormat);
        put(kind);
    }
    bool check_header(uint32_t kind) {
        uint32_t m = 0, v = 0, k = 0;
        get(m), get(v), get(k);
        return ok && m == kMagic && v == kFormat && k == kind;
    }
};

std::string path_of(const char *what, const CacheKey &key, const std::string &tag) {
    std::string p = cache_dir() + "/" + what + "_" + key.hex();
    if (!tag.empty()) {
        Hasher h(7);
        h.bytes(tag.data(), tag.size());
        char buf[24];
        snprintf(buf, sizeof buf, "_%016llx", (unsigned long long) h.h);
        p += buf;
    }
    return p + ".chb";
}

void io_panel(File &f, SweepImage &img, bool write) {
    if (write) {
        f.put(img.upper), f.put(img.elem), f.put(img.n), f.put(img.page_bytes), f.put(img.n_slots), f.put(img.n_contrib_rows);
        f.put(img.max_local_rows), f.put(img.n_levels), f.put(img.n_values), f.put(img.n_stored_values), f.put(img.n_blocks);
        f.put_vec(img.slab), f.put_vec(img.tasks), f.put_vec(img.order);
    } else {
        f.get(img.upper), f.get(img.elem), f.get(img.n), f.get(img.page_bytes), f.get(img.n_slots), f.get(img.n_contrib_rows);
        f.get(img.max_local_rows), f.get(img.n_levels), f.get(img.n_values), f.get(img.n_stored_values), f.get(img.n_blocks);
        f.get_vec(img.slab), f.get_vec(img.tasks), f.get_vec(img.order);
    }
}
void io_front(File &f, FrontImage &img, bool write) {
    if (write) {
        f.put(img.upper), f.put(img.elem), f.put(img.n), f.put(img.tile_rows), f.put(img.n_warps), f.put(img.page_bytes);
        f.put(img.n_w_rows), f.put(img.max_stage_rows), f.put(img.n_levels), f.put(img.n_values), f.put(img.n_stored_values);
        f.put(img.n_local_tasks), f.put(img.n_dense_tasks), f.put(img.n_mma), f.put(img.n_asm_entries);
        f.put_vec(img.slab), f.put_vec(img.tasks), f.put_vec(img.order);
    } else {
        f.get(img.upper), f.get(img.elem), f.get(img.n), f.get(img.tile_rows), f.get(img.n_warps), f.get(img.page_bytes);
        f.get(img.n_w_rows), f.get(img.max_stage_rows), f.get(img.n_levels), f.get(img.n_values), f.get(img.n_stored_values);
        f.get(img.n_local_tasks), f.get(img.n_dense_tasks), f.get(img.n_mma), f.get(img.n_asm_entries);
        f.get_vec(img.slab), f.get_vec(img.tasks), f.get_vec(img.order);
    }
}

}  // namespace

std::string CacheKey::hex() const {
    char buf[40];
    snprintf(buf, sizeof buf, "%016llx%016llx", (unsigned long long) a, (unsigned long long) b);
    return buf;
}

std::string cache_dir() {
    const char *d = std::getenv("CHB_CACHE_DIR");
    if (!d || !*d) return "";
    struct stat st;
    if (stat(d, &st) != 0) mkdir(d, 0777);
    if (stat(d, &st) != 0 || !S_ISDIR(st.st_mode)) return "";
    return d;
}

CacheKey matrix_key(const CscMatrix &A) { return hash_arrays(A.n, A.colptr, A.rowidx, A.vals, 1); }
CacheKey factor_key(const HostFactor &L) {
    CacheKey k = hash_arrays(L.n, L.colptr, L.rowidx, L.vals, 2);
    Hasher h(k.a);
    h.bytes(L.perm.data(), sizeof(int) * L.perm.size());
    k.a = h.h;
    return k;
}

bool load_factor(const CacheKey &key, HostFactor *out) {
    if (cache_dir().empty()) return false;
    File f;
    if (!f.open_read(path_of("factor", key, "")) || !f.check_header(1)) return false;
    HostFactor L;
    f.get(L.n);
    f.get_vec(L.perm), f.get_vec(L.colptr), f.get_vec(L.rowidx), f.get_vec(L.vals);
    if (!f.ok || L.n < 0 || (int) L.perm.size() != L.n || (int) L.colptr.size() != L.n + 1 ||
        L.rowidx.size() != L.vals.size() || (L.n > 0 && (size_t) L.colptr[L.n] != L.rowidx.size()))
        return false;
    *out = std::move(L);
    return true;
}

void store_factor(const CacheKey &key, const HostFactor &L) {
    if (cache_dir().empty()) return;
    File f;
    if (!f.open_write(path_of("factor", key, ""))) return;
    f.put_header(1);
    f.put(L.n);
    f.put_vec(L.perm), f.put_vec(L.colptr), f.put_vec(L.rowidx), f.put_vec(L.vals);
    f.commit();
}

bool load_panel_images(const CacheKey &key, const std::string &tag, SweepImage *lower, SweepImage *upper) {
    if (cache_dir().empty()) return false;
    File f;
    if (!f.open_read(path_of("panel", key, tag)) || !f.check_header(2)) return false;
    io_panel(f, *lower, false);
    io_panel(f, *upper, false);
    return f.ok && !lower->upper && upper->upper;
}
void store_panel_images(const CacheKey &key, const std::string &tag, const SweepImage &lower, const SweepImage &upper) {
    if (cache_dir().empty()) return;
    File f;
    if (!f.open_write(path_of("panel", key, tag))) return;
    f.put_header(2);
    io_panel(f, const_cast<SweepImage &>(lower), true);
    io_panel(f, const_cast<SweepImage &>(upper), true);
    f.commit();
}
bool load_front_images(const CacheKey &key, const std::string &tag, FrontImage *lower, FrontImage *upper) {
    if (cache_dir().empty()) return false;
    File f;
    if (!f.open_read(path_of("front", key, tag)) || !f.check_header(3)) return false;
    io_front(f, *lower, false);
    io_front(f, *upper, false);
    return f.ok && !lower->upper && upper->upper;
}
void store_front_images(const CacheKey &key, const std::string &tag, const FrontImage &lower, const FrontImage &upper) {
    if (cache_dir().empty()) return;
    File f;
    if (!f.open_write(path_of("front", key, tag))) return;
    f.put_header(3);
    io_front(f, const_cast<FrontImage &>(lower), true);
    io_front(f, const_cast<FrontImage &>(upper), true);
    f.commit();
}

}  // namespace chb
