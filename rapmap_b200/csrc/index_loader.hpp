// Host-side parser of the on-disk RapMapSAIndex (quasiindex output), loaded UNCHANGED:
//   header.json, sa.bin, txpInfo.bin, rsd.bin, hash.bin (dense) | hash_info.bph + hash_info.val (-p).
// Replaces RapMapSAIndex<IndexT,HashT>::load (reference src/RapMapSAIndex.cpp:96-176) without cereal /
// sparsepp: the formats are plain little-endian PODs (writers: src/RapMapSAIndexer.cpp:109-110,:694-733,
// :791-818; sparsepp layout include/sparsepp/spp.h:2355-2366,:1603-1613 and
// include/SparseHashSerializer.hpp:29-47).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rapmap_b200 {

struct KmerRecord {  // one dense hash.bin record for a 32-bit index
  uint64_t kmer;
  int32_t begin, end;
};

struct HostIndex {
  uint32_t k{31};
  bool bigSA{false};
  bool perfectHash{false};
  std::vector<int32_t> SA;            // suffix array (32-bit indexes only, see load())
  std::string text;                   // concatenated transcripts, '$' after each
  std::vector<std::string> txpNames;
  std::vector<int32_t> txpOffsets;
  std::vector<int32_t> txpLens;       // derived exactly as src/RapMapSAIndex.cpp:151-163
  std::vector<uint32_t> txpCompleteLens;
  uint64_t numBits{0};
  std::vector<uint64_t> rsdBits;      // '$' positions
  std::vector<KmerRecord> kmers;      // dense: parsed from hash.bin; perfect: rebuilt from SA+text (see .cpp)
  // Returns false and fills err on failure.
  bool load(const std::string& dir, std::string& err);
};

} // namespace rapmap_b200
