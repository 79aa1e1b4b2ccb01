// =================================================================================================
// CPU ORACLE — TEST INFRASTRUCTURE ONLY.
//
// A plain, sequential C++ restatement of the reference's `rapmap quasimap` per-read hot path
// (COMBINE-lab/RapMap v0.6.0).  It exists so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg can CHECK the CUDA path; nothing under rapmap_b200/ links, imports or calls it.
//
// Parity status: PINNED.  The reference's own tests hold no golden vectors for this path
// (SURVEY.md §4), so the oracle is pinned against outputs of the reference itself, compiled
// unmodified into oracle/_ref/rapmap_ref by oracle/build_ref.sh: byte-identical SAM on
// sample_data (md5 5271acf4... / ddd30824... with -s) and on synthetic read sets
// (tests/test_oracle_vs_reference.py; fixtures under tests/golden/).
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// =================================================================================================
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace oracle {

// ---- index (include/RapMapSAIndex.hpp:46-83, src/RapMapSAIndex.cpp:96-176) ----------------------
struct Index {
  int k{31};
  bool bigSA{false}, perfectHash{false};
  std::vector<int64_t> SA;
  std::string seq;                       // concatenated text, '$' after every transcript
  std::vector<std::string> txpNames;
  std::vector<int64_t> txpOffsets, txpLens;
  std::vector<uint32_t> txpCompleteLens;
  std::vector<uint64_t> bits;            // rsd.bin
  std::vector<uint64_t> rankBlock;       // cumulative popcount before each 64-bit word
  std::unordered_map<uint64_t, std::pair<int64_t, int64_t>> khash;  // k-mer -> [begin,end)
  bool load(const std::string& dir, std::string* err);
  // src/RapMapSAIndex.cpp:91-94 -> src/rank9b.cpp:55-60 : #set bits strictly before p
  int64_t transcriptAtPosition(int64_t p) const;
  const std::pair<int64_t, int64_t>* find(uint64_t kmer) const {
    auto it = khash.find(kmer);
    return it == khash.end() ? nullptr : &it->second;
  }
};

// ---- options (src/RapMapSAMapper.cpp:114-152; derived values :1113-1135, :385-455) --------------
struct Opts {
  uint32_t maxNumHits{200};
  double quasiCov{0.0};
  bool sensitive{true};        // = !--noSensitive  (true => NIP disabled)
  bool strictCheck{true};      // = !--noStrictCheck
  bool fuzzy{false};
  bool selAln{false};
  float consensusSlack{0.0f};  // 0.2 when selAln unless overridden
  double minScoreFraction{0.65};
  int matchScore{2}, mismatchPenalty{-4}, gapOpenPenalty{4}, gapExtendPenalty{2};
  int dpBandwidth{15};
  bool hardFilter{false};
  int alignmentPolicy{0};      // 0 DEFAULT, 1 BT2, 2 BT2_STRICT
  bool noOrphans{false}, noDovetail{false};
  int maxMMPExtension{7};
  bool recoverOrphans{false};  // --recoverOrphans (needs -s or -f)
};

enum MateStatus : uint8_t { SINGLE_END = 0, PAIRED_END_LEFT = 1, PAIRED_END_RIGHT = 2, PAIRED_END_PAIRED = 3 };
enum ChainStatus : uint8_t { PERFECT = 0, UNGAPPED = 1, ALIGNED_ON_LEFT = 2, ALIGNED_ON_RIGHT = 3, REGULAR = 4 };

// include/RapMapUtils.hpp:516-525
struct SAIntervalHit {
  int64_t begin, end;
  uint32_t len, queryPos;
  bool queryRC;
  int64_t span() const { return end - begin; }
};
// include/HitManager.hpp:59-72
struct HitCollectorInfo {
  size_t readLen{0};
  int32_t maxDist{0};
  std::vector<SAIntervalHit> fwdSAInts, rcSAInts;
};

// include/RapMapUtils.hpp:399-502 (fields that reach the output or a later decision)
struct QuasiAlignment {
  uint32_t tid{0};
  int32_t pos{0}, matePos{0};
  bool fwd{true}, mateIsFwd{true};
  uint32_t fragLen{0}, readLen{0}, mateLen{0};
  bool isPaired{false};
  uint8_t mateStatus{SINGLE_END};
  double score{1.0};
  int32_t alnScore{0};
  uint8_t chainLeft{REGULAR}, chainRight{REGULAR};
  double chainScore;
  bool hasMultiPos{false};
  std::vector<int32_t> allPositions, oppositeStrandPositions;
  QuasiAlignment();
};

struct Counters { uint64_t numReads{0}, peHits{0}, seHits{0}, totHits{0}, tooManyHits{0}; };

// Operation counters for the roofline's "algorithmic bytes per pair" (SURVEY.md §8d).
struct OpCounts { uint64_t hashFind{0}, saProbes{0}, textCmp{0}, rankCalls{0}, intervals{0}, kswCalls{0}, alnCalls{0}, kswCells{0}; };

class Mapper {
public:
  Mapper(const Index& idx, const Opts& o);
  // include/SACollector.hpp:108-362
  bool collect(const std::string& read, HitCollectorInfo& hc);
  // src/HitManager.cpp:691-882
  void hitsToMappingsSimple(uint8_t mateStatus, HitCollectorInfo& hc, std::vector<QuasiAlignment>& hits);
  // src/RapMapSAMapper.cpp:461-711 (one pair); jointHits is the final per-pair result
  void mapPair(const std::string& r1, const std::string& r2, std::vector<QuasiAlignment>& jointHits);
  // src/RapMapSAMapper.cpp:156-371 (one unmated read)
  void mapSingle(const std::string& r, std::vector<QuasiAlignment>& hits);
  // src/RapMapUtils.cpp:313-588 / :137-196 (SAM text for one pair)
  void samPair(const std::string& n1, const std::string& s1, const std::string& n2, const std::string& s2,
               std::vector<QuasiAlignment>& jointHits, std::string& out);
  void samSingle(const std::string& name, const std::string& seq, std::vector<QuasiAlignment>& hits, std::string& out);
  std::string samHeader() const;
  // src/ksw2pp/KSW2Aligner.cpp:205-234 -> src/ksw2pp/ksw2_extz2_sse.c:18-304 (score only); returns max(mqe,mte)
  int32_t kswExtzScore(const char* q, int qlen, const char* t, int tlen);
  Counters ctr;
  OpCounts ops;
  const Index& idx;
  Opts o;

private:
  struct Impl;
  // derived per-thread settings (src/RapMapSAMapper.cpp:385-455)
  bool disableNIP_, strictCheck_, doChaining_, considerMultiPos_;
  double covReq_;
  int64_t maxInterval_{1000};
  int32_t maxMMPExtension_{7}, strictCheckSlack_{0};
  float consensusFraction_{1.0f};
  friend struct Impl;
};

// include/Kmer.hpp:524-542 — returns false at the first non-ACGT char, leaving the partial word
bool encodeKmer(const char* s, int k, uint64_t& w);
// include/Kmer.hpp:92-100
uint64_t kmerRC(uint64_t w, int k);
// src/RapMapUtils.cpp:107-128
void reverseRead(const std::string& seq, std::string& out);

} // namespace oracle
