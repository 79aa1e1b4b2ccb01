// k-mer and base primitives shared by the kernels: Kmer<32,1> reverse complement / homopolymer test
// (reference include/Kmer.hpp:92-100,484-487) and the reverseRead table (src/RapMapUtils.cpp:63-72).
#pragma once
#include "kernels.cuh"

namespace rapmap_b200 {

__device__ __forceinline__ uint64_t kmerRC(uint64_t w, int k) {  // include/Kmer.hpp:92-100
  // reverse the 2-bit groups: full bit reversal (two BREVs), then swap the two bits of every group back
  const uint32_t lo = __brev(static_cast<uint32_t>(w >> 32)), hi = __brev(static_cast<uint32_t>(w));
  const uint32_t lo2 = ((lo & 0x55555555u) << 1) | ((lo >> 1) & 0x55555555u), hi2 = ((hi & 0x55555555u) << 1) | ((hi >> 1) & 0x55555555u);
  const uint64_t r = (static_cast<uint64_t>(hi2) << 32) | lo2;
  return (~r) >> (2 * (32 - k));
}

__device__ __forceinline__ bool isHomopolymer(uint64_t w, int k) {  // include/Kmer.hpp:484-487
  uint64_t mask = (1ULL << (2 * k)) - 1ULL;
  uint64_t nuc = w & 3ULL;
  return w == (mask & ((w << 2) | nuc));
}

__device__ __forceinline__ uint8_t upperChar(uint8_t c) { return (c >= 'a' && c <= 'z') ? static_cast<uint8_t>(c - 32) : c; }

__device__ __forceinline__ uint8_t rcChar(uint8_t c) {  // rapmap::utils::reverseRead table, src/RapMapUtils.cpp:63-72
  switch (c | 0x20) {
    case 'a': return 'T';
    case 'c': return 'G';
    case 'g': return 'C';
    case 't': return 'A';
    case 'u': return (c == 'U' || c == 'u') ? 'A' : 'N';
    default: return 'N';
  }
}

} // namespace rapmap_b200
