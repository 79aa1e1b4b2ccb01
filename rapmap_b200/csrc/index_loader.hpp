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

// Perfect-hash flavour (-p): boomphf::mphf levels (reference include/BooPHF.hpp:1172-1245) + FrugalBooMap values
// (include/FrugalBooMap.hpp:199-247), parsed from hash_info.bph / hash_info.val unchanged.
struct PhfLevel {
  uint64_t sizeBits{0};            // bitVector::_size
  uint64_t hashDomain{0};          // recomputed at load exactly as mphf::load does (pow on the host)
  std::vector<uint64_t> bits;      // _nchar words
  std::vector<uint64_t> ranks;     // one sample per 512 bits, offset by the keys of the previous levels
};
struct HostPhf {
  double gamma{0};
  int32_t nbLevels{0};
  uint64_t lastBitsetRank{0}, nelem{0};
  std::vector<PhfLevel> levels;
  std::vector<std::pair<uint64_t, uint64_t>> finalHash;   // sorted by key
  std::vector<int32_t> data;                               // interval start per MPHF slot
  std::vector<uint8_t> lens;                               // interval length, 255 => overflow
  std::vector<std::pair<int32_t, int32_t>> overflow;       // sorted by start: start -> length
};

struct HostIndex {
  uint32_t k{31};
  bool bigSA{false};
  bool perfectHash{false};
  bool unsupported{false};            // load() failed because of a flavour the device path does not implement (not a malformed file)
  std::vector<int32_t> SA;            // suffix array (32-bit indexes only, see load())
  std::string text;                   // concatenated transcripts, '$' after each
  std::vector<std::string> txpNames;
  std::vector<int32_t> txpOffsets;
  std::vector<int32_t> txpLens;       // derived exactly as src/RapMapSAIndex.cpp:151-163
  std::vector<uint32_t> txpCompleteLens;
  uint64_t numBits{0};
  std::vector<uint64_t> rsdBits;      // '$' positions
  std::vector<KmerRecord> kmers;      // dense index: records of hash.bin
  HostPhf phf;                        // -p index: BooPHF + FrugalBooMap arrays
  // Returns false and fills err on failure.
  bool load(const std::string& dir, std::string& err);
};

} // namespace rapmap_b200
