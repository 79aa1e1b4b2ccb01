// Device image of a RapMapSAIndex and the primitive lookups the kernels share.
//
// HBM layout (one contiguous blob, sections 256-byte aligned, so the whole index is a single
// ncclBroadcast / cudaMemcpy):
//   ImageHeader | SA int32[n] | text u8[n + pad] | rank uint4[ceil(n/64)+1] | txpOffsets int32[T] |
//   txpLens int32[T] | hash table uint4[slots]
// * rank: one 16-byte record per 64 text positions {bits_lo, bits_hi, #'$' before this word, 0}: a
//   transcript id is ONE 16-byte load + popcount (replaces rank9b::rank, reference src/rank9b.cpp:55-60,
//   which needs three loads; same value: number of set bits strictly before p).
// * hash table: open addressing, linear probing, 16-byte slots {kmer_lo, kmer_hi, begin, end}, capacity a
//   power of two >= 2 x #k-mers; two slots share a 32-byte DRAM sector.  Replaces the sparsepp
//   RegHashT<uint64_t, SAInterval> (reference include/RapMapUtils.hpp:65-67): same key -> [begin,end) map.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rapmap_b200 {

struct ImageHeader {
  uint64_t magic;          // 'RMB2IMG1'
  uint64_t totalBytes;
  uint64_t n;              // text / SA length
  uint64_t numTxp;
  uint64_t tableSlots;     // power of two
  uint64_t numKmers;
  uint32_t k;
  uint32_t pad;
  uint64_t offSA, offText, offRank, offTxpOffsets, offTxpLens, offTable;
};
static constexpr uint64_t kImageMagic = 0x31474D4932424D52ULL;

struct DeviceIndex {
  const int32_t* SA;
  const uint8_t* text;
  const uint4* rank;
  const int32_t* txpOffsets;
  const int32_t* txpLens;
  const uint4* table;
  uint64_t tableMask;
  int64_t n;
  uint32_t k;
  uint32_t numTxp;
};

static constexpr uint64_t kEmptyKey = ~0ULL;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

#ifdef __CUDACC__
// k-mer -> SA interval; {-1,-1} when absent.  (RegHashT::find, reference include/SACollector.hpp:196,541)
__device__ __forceinline__ int2 hashFind(const DeviceIndex& ix, uint64_t key) {
  uint64_t s = mix64(key) & ix.tableMask;
  while (true) {
    uint4 e = __ldg(ix.table + s);
    uint64_t kk = (static_cast<uint64_t>(e.y) << 32) | e.x;
    if (kk == key) return make_int2(static_cast<int>(e.z), static_cast<int>(e.w));
    if (kk == kEmptyKey) return make_int2(-1, -1);
    s = (s + 1) & ix.tableMask;
  }
}

// RapMapSAIndex::transcriptAtPosition (reference src/RapMapSAIndex.cpp:91-94): #'$' strictly before p.
__device__ __forceinline__ uint32_t transcriptAt(const DeviceIndex& ix, int64_t p) {
  uint4 r = __ldg(ix.rank + (p >> 6));
  uint64_t bits = (static_cast<uint64_t>(r.y) << 32) | r.x;
  uint64_t m = (1ULL << (p & 63)) - 1ULL;
  return r.z + __popcll(bits & m);
}
#endif

} // namespace rapmap_b200
