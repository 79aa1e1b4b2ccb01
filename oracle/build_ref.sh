#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (see oracle/README.md).
# Compiles the UNMODIFIED reference (COMBINE-lab/RapMap v0.6.0) from the sources where they lie
# under /root/reference into oracle/_ref/rapmap_ref, without running the reference's own build
# system: plain gcc/g++ on the source list of src/CMakeLists.txt:1-15,70-84, libdivsufsort compiled
# from the vendored tarball (headers generated with sed instead of cmake's configure_file), and
# the cereal stand-in of oracle/cereal_standin/ in place of the un-vendored cereal v1.2.2.
# Nothing is copied into the repository; all outputs go to oracle/_ref/ (git-ignored).
# Flags: -O3 -std=c++14, no -march=native, -ffp-contract=off (pins the chain-score float maths,
# SURVEY.md §7.3-2).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${RAPMAP_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
BLD="$OUT/build"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then
  echo "[build_ref] $REF not present; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$BLD/obj" "$BLD/dss/include"

# ---- libdivsufsort (suffix-array construction, index time only) --------------------------------
if [ ! -f "$BLD/libdivsufsort.a" ] || [ ! -f "$BLD/libdivsufsort64.a" ]; then
  tar xzf "$REF/external/libdivsufsort.tar.gz" -C "$BLD/dss"
  DSS="$BLD/dss/libdivsufsort-master"
  gen_header() { # $1 = "" | "64", $2 = index type, $3 = PRId
    sed -e "s/@W64BIT@/$1/g" -e "s/@INCFILE@/#include <inttypes.h>/" \
        -e "s/@DIVSUFSORT_EXPORT@//" -e "s/@DIVSUFSORT_IMPORT@//" \
        -e "s/@SAUCHAR_TYPE@/uint8_t/" -e "s/@SAINT32_TYPE@/int32_t/" \
        -e "s/@SAINDEX_TYPE@/$2/" -e "s/@SAINT_PRId@/PRId32/" -e "s/@SAINDEX_PRId@/$3/" \
        "$DSS/include/divsufsort.h.cmake" > "$BLD/dss/include/divsufsort$1.h"
  }
  gen_header "" int32_t PRId32
  gen_header 64 int64_t PRId64
  cat > "$BLD/dss/include/config.h" <<'EOC'
#ifndef _CONFIG_H
#define _CONFIG_H 1
#define PROJECT_VERSION_FULL "2.0.2-1"
#define HAVE_INTTYPES_H 1
#define HAVE_STDDEF_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_STRINGS_H 1
#define HAVE_MEMORY_H 1
#define HAVE_SYS_TYPES_H 1
#ifndef INLINE
# define INLINE inline
#endif
#endif
EOC
  for v in "" 64; do
    objs=()
    for f in divsufsort sssort trsort utils; do
      def=""; [ "$v" = 64 ] && def="-DBUILD_DIVSUFSORT64"
      gcc -O3 -fomit-frame-pointer -fopenmp -DHAVE_CONFIG_H $def -I"$BLD/dss/include" -I"$DSS/include" \
          -c "$DSS/lib/$f.c" -o "$BLD/obj/dss${v}_$f.o"
      objs+=("$BLD/obj/dss${v}_$f.o")
    done
    rm -f "$BLD/libdivsufsort$v.a"
    ar rcs "$BLD/libdivsufsort$v.a" "${objs[@]}"
  done
fi

# ---- reference sources -------------------------------------------------------------------------
INC=(-I"$REF/include" -I"$HERE/cereal_standin" -I"$BLD/dss/include" -I"$REF/external")
CXXF=(-O3 -std=c++14 -pthread -ffp-contract=off -w -DHAVE_SIMDE=0 -DRAPMAP_SALMON_SUPPORT=0)
CXXF=(-O3 -std=c++14 -pthread -ffp-contract=off -w)
CF=(-O3 -pthread -ffp-contract=off -w)
compile() { # $1 = compiler, $2 = src, $3 = obj, rest = flags
  local cc="$1" src="$2" obj="$3"; shift 3
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then echo "  CC $(basename "$src")"; "$cc" "$@" -c "$src" -o "$obj"; fi
}
pids=()
run() { "$@" & pids+=($!); if [ "${#pids[@]}" -ge "$JOBS" ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi; }
for f in RapMap RapMapSAIndexer RapMapUtils RapMapSAMapper RapMapFileSystem RapMapSAIndex HitManager FastxParser rank9b edlib; do
  run compile g++ "$REF/src/$f.cpp" "$BLD/obj/$f.o" "${CXXF[@]}" "${INC[@]}"
done
run compile g++ "$REF/src/stringpiece.cc" "$BLD/obj/stringpiece.o" "${CXXF[@]}" "${INC[@]}"
run compile g++ "$REF/src/metro/metrohash64.cpp" "$BLD/obj/metrohash64.o" "${CXXF[@]}" "${INC[@]}"
run compile gcc "$REF/src/xxhash.c" "$BLD/obj/xxhash.o" "${CF[@]}" "${INC[@]}"
run compile gcc "$REF/src/bit_array.c" "$BLD/obj/bit_array.o" "${CF[@]}" "${INC[@]}"
KD=(-DKSW_CPU_DISPATCH -DHAVE_KALLOC)
for f in kalloc ksw2_extd ksw2_extz ksw2_gg ksw2_gg2 ksw2_gg2_sse; do
  run compile gcc "$REF/src/ksw2pp/$f.c" "$BLD/obj/$f.o" "${CF[@]}" "${KD[@]}" "${INC[@]}"
done
run compile g++ "$REF/src/ksw2pp/KSW2Aligner.cpp" "$BLD/obj/KSW2Aligner.o" "${CXXF[@]}" "${KD[@]}" "${INC[@]}"
for f in ksw2_extd2_sse ksw2_extf2_sse ksw2_extz2_sse; do
  run compile gcc "$REF/src/ksw2pp/$f.c" "$BLD/obj/${f}_sse2.o" "${CF[@]}" -msse2 -mno-sse4.1 "${KD[@]}" -DKSW_SSE2_ONLY "${INC[@]}"
  run compile gcc "$REF/src/ksw2pp/$f.c" "$BLD/obj/${f}_sse41.o" "${CF[@]}" -msse4.1 "${KD[@]}" "${INC[@]}"
done
for p in "${pids[@]}"; do wait "$p"; done
MAIN_OBJS=()
for f in RapMap RapMapSAIndexer RapMapUtils RapMapSAMapper RapMapFileSystem RapMapSAIndex HitManager FastxParser rank9b edlib \
         stringpiece metrohash64 xxhash bit_array kalloc ksw2_extd ksw2_extz ksw2_gg ksw2_gg2 ksw2_gg2_sse KSW2Aligner \
         ksw2_extd2_sse_sse2 ksw2_extf2_sse_sse2 ksw2_extz2_sse_sse2 ksw2_extd2_sse_sse41 ksw2_extf2_sse_sse41 ksw2_extz2_sse_sse41; do
  MAIN_OBJS+=("$BLD/obj/$f.o")
done
g++ -O3 -pthread -o "$OUT/rapmap_ref" "${MAIN_OBJS[@]}" "$BLD/libdivsufsort.a" "$BLD/libdivsufsort64.a" -lz -lm -lgomp -lrt -ldl
echo "[build_ref] built $OUT/rapmap_ref"

# ---- stage-dump harness (reference header templates called directly) ---------------------------
if [ -f "$HERE/ref_stage_dump.cpp" ]; then
  g++ "${CXXF[@]}" "${INC[@]}" -o "$OUT/ref_stage_dump" "$HERE/ref_stage_dump.cpp" \
      "$BLD/obj/RapMapSAIndex.o" "$BLD/obj/HitManager.o" "$BLD/obj/RapMapUtils.o" "$BLD/obj/rank9b.o" \
      "$BLD/obj/bit_array.o" "$BLD/obj/xxhash.o" "$BLD/obj/FastxParser.o" "$BLD/obj/stringpiece.o" -lz -lm -pthread
  echo "[build_ref] built $OUT/ref_stage_dump"
fi
