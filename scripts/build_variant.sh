#!/bin/bash
# build an A/B variant of the library: scripts/build_variant.sh NAME -DFLAG=..   -> rapmap_b200/_build/ab/lib_NAME.so
name=$1; shift
mkdir -p rapmap_b200/_build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Wno-deprecated-declarations \
  -Xptxas -v "$@" -shared -o rapmap_b200/_build/ab/lib_$name.so rapmap_b200/csrc/capi.cu rapmap_b200/csrc/index_loader.cpp 2>&1 | grep -A2 "${GREP:-sa_collect_lane_kernel}" | grep -E "Used|spill" | tr '\n' ' '
echo " <- $name"
