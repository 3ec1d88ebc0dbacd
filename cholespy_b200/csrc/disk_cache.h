// Library-level disk cache of the host-side products of a constructor: the CHOLMOD factor (minutes for the BASELINE
// meshes: 105 s for cfg4) and the sweep images built from it (seconds), keyed by a hash of the canonical CSC matrix.
//
// New: the reference re-runs CHOLMOD in every constructor (src/cholesky_solver.cpp:200-242); SURVEY.md 8(f)-2 ranks a
// factor/schedule cache second among the follow-ups.  Opt-in: the cache lives in the directory named by the
// environment variable CHB_CACHE_DIR and does nothing when it is unset.  Files are written to a temporary name and
// renamed, so concurrent constructors (one process per GPU) never see a partial file.
#pragma once
#include <cstdint>
#include <string>

#include "factor_host.h"
#include "front_schedule.h"
#include "panel_schedule.h"

namespace chb {

struct CacheKey {
    uint64_t a = 0, b = 0;
    std::string hex() const;
};

std::string cache_dir();                      // "" when caching is off
CacheKey matrix_key(const CscMatrix &A);      // 128-bit hash of n, colptr, rowidx and the values' bit patterns
CacheKey factor_key(const HostFactor &L);     // ... of a factor handed in directly (chb_solver_create_from_factor)

bool load_factor(const CacheKey &key, HostFactor *out);
void store_factor(const CacheKey &key, const HostFactor &L);

// `tag`: everything else the images depend on (element size, schedule options, format version)
bool load_panel_images(const CacheKey &key, const std::string &tag, SweepImage *lower, SweepImage *upper);
void store_panel_images(const CacheKey &key, const std::string &tag, const SweepImage &lower, const SweepImage &upper);
bool load_front_images(const CacheKey &key, const std::string &tag, FrontImage *lower, FrontImage *upper);
void store_front_images(const CacheKey &key, const std::string &tag, const FrontImage &lower, const FrontImage &upper);

}  // namespace chb
