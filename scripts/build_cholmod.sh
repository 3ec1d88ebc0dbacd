#!/usr/bin/env bash
# Build the host factorisation dependency (CHOLMOD 3.0.13 / SuiteSparse 5.4.0 / METIS 5.1.0)
# from the reference's vendored submodule, compiling the sources WHERE THEY LIE under
# $CHOLESPY_REF (default /root/reference). Nothing is copied into this repository; the only
# output is cholespy_b200/lib/libcholmod_b200.so (git-ignored, shipped to the GPU box).
#
# CHOLMOD's host factorisation is kept exactly as the reference uses it
# (src/cholesky_solver.cpp:200-242: simplicial LL', NESDIS); it is not on the hot path.
# -DNSUPERNODAL drops the BLAS/LAPACK-dependent supernodal module, which the reference never
# reaches (it forces CHOLMOD_SIMPLICIAL, src/cholesky_solver.cpp:206).
set -euo pipefail
REF=${CHOLESPY_REF:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=$HERE/cholespy_b200/lib/libcholmod_b200.so
SS=$REF/ext/suitesparse-metis-for-windows/SuiteSparse
if [ ! -d "$SS" ]; then
  if [ -f "$OUT" ]; then echo "build_cholmod: reference tree absent, keeping prebuilt $OUT"; exit 0; fi
  echo "build_cholmod: $SS not found and no prebuilt library" >&2; exit 1
fi
if [ -f "$OUT" ] && [ "${FORCE:-0}" != "1" ]; then echo "build_cholmod: $OUT up to date"; exit 0; fi
BUILD=${BUILD_DIR:-/tmp/cholmod_b200_build}
mkdir -p "$BUILD"
INC="-I$SS -I$SS/SuiteSparse_config -I$SS/AMD/Include -I$SS/CAMD/Include -I$SS/COLAMD/Include \
 -I$SS/CCOLAMD/Include -I$SS/CHOLMOD/Include -I$SS/metis-5.1.0/include \
 -I$SS/metis-5.1.0/GKlib -I$SS/metis-5.1.0/libmetis"
CFLAGS="-O2 -fPIC -w -DNCOMPLEX -DNSUPERNODAL -D_GNU_SOURCE -DNDEBUG"
list=$BUILD/srcs.txt; : > "$list"
echo "$SS/SuiteSparse_config/SuiteSparse_config.c cfg" >> "$list"
for lib in AMD CAMD COLAMD CCOLAMD CHOLMOD; do
  for f in "$SS/$lib/SourceWrappers/"*.c; do
    case "$f" in *super_*) continue;; esac
    echo "$f $lib" >> "$list"
  done
done
for f in "$SS/metis-5.1.0/GKlib/"*.c;    do echo "$f gk"    >> "$list"; done
for f in "$SS/metis-5.1.0/libmetis/"*.c; do echo "$f metis" >> "$list"; done
# unique object names: <tag>__<basename>.o
awk '{n=split($1,a,"/"); b=a[n]; gsub(/\.c$/,"",b); gsub(/\./,"_",b); print $1, $2"__"b".o"}' "$list" > "$BUILD/jobs.txt"
export INC CFLAGS BUILD
xargs -P "$(nproc)" -L 1 sh -c 'gcc $CFLAGS $INC -c "$0" -o "$BUILD/$1"' < "$BUILD/jobs.txt"
gcc -shared -o "$OUT" "$BUILD"/*.o -lm
echo "build_cholmod: wrote $OUT ($(ls "$BUILD"/*.o | wc -l) objects)"
