#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/adapter_sam: tests/cpp/adapter_sam.cpp (the stub of INTEGRATION.md §1) compiled
# against the UNMODIFIED reference headers under /root/reference, linked with the reference objects build_ref.sh left in
# oracle/_ref/build/obj and with librapmap_cuda.so.  It proves that include/rapmap_b200/adapter.hpp drops into the
# reference's own source tree; tests/test_gpu_adapter.py runs it on the GPU box (the binary travels, the sources do not).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${RAPMAP_REFERENCE_DIR:-/root/reference}"
BLD="$HERE/_ref/build"
if [ ! -d "$REF/include" ] || [ ! -f "$BLD/obj/RapMapUtils.o" ]; then
  echo "[adapter_harness] reference sources / objects not present; keeping prebuilt oracle/_ref/adapter_sam as is" >&2
  exit 0
fi
g++ -O2 -std=c++14 -pthread -w -I"$REF/include" -I"$HERE/cereal_standin" -I"$REF/external" -I"$ROOT/include" \
    -o "$HERE/_ref/adapter_sam" "$ROOT/tests/cpp/adapter_sam.cpp" \
    "$BLD/obj/RapMapSAIndex.o" "$BLD/obj/HitManager.o" "$BLD/obj/RapMapUtils.o" "$BLD/obj/rank9b.o" "$BLD/obj/bit_array.o" \
    "$BLD/obj/xxhash.o" "$BLD/obj/FastxParser.o" "$BLD/obj/stringpiece.o" \
    -L"$ROOT/rapmap_b200/_build" -lrapmap_cuda '-Wl,-rpath,$ORIGIN/../../rapmap_b200/_build' -lz -lm
echo "[adapter_harness] built $HERE/_ref/adapter_sam"
