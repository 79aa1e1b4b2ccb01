// Source-level adapter between the reference's C++ types and the C-ABI of librapmap_cuda.so.
//
// Salmon and `rapmap quasimap` consume the mapping path as headers: per read SACollector::operator()
// (include/SACollector.hpp:108), hit_manager::hitsToMappingsSimple (src/HitManager.cpp:691),
// rapmap::utils::mergeLeftRightHits / mergeLeftRightHitsFuzzy (include/RapMapUtils.hpp:1185, :864) and
// selective_alignment::utils::getAlnScore (include/SelectiveAlignmentUtils.hpp:260), collecting
// std::vector<rapmap::utils::QuasiAlignment> per read pair.  A per-read call cannot feed a GPU, so the adapter offers the
// SAME result type one level up: BatchMapper maps a whole fastx_parser::ReadGroup chunk with one rapmap_cuda_map_batch call
// and hands back one std::vector<QuasiAlignment> per pair, ready for writeAlignmentsToStream / Salmon's jointHitGroup.
//
// Header-only, C++14.  It includes the CONSUMER's own RapMap headers (RapMapUtils.hpp, FastxParser.hpp): compile it inside
// the RapMap / Salmon tree with -I<this repo>/include and link -lrapmap_cuda (CMake: target rapmap_b200::rapmap_cuda).
// tests/cpp/adapter_sam.cpp is the stub of INTEGRATION.md §1 built this way against the unmodified reference headers; its
// SAM output through the reference's own writeAlignmentsToStream is md5-checked against the golden files.
#ifndef RAPMAP_B200_ADAPTER_HPP
#define RAPMAP_B200_ADAPTER_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "FastxParser.hpp"
#include "RapMapUtils.hpp"
#include "rapmap_cuda.h"

namespace rapmap_b200 {

// rapmap_hit_t -> QuasiAlignment (include/RapMapUtils.hpp:399-502): the fields that reach SAM and Salmon.
inline rapmap::utils::QuasiAlignment toQuasiAlignment(const rapmap_hit_t& h) {
  using rapmap::utils::ChainStatus;
  using rapmap::utils::MateStatus;
  rapmap::utils::QuasiAlignment qa(h.tid, h.pos, h.fwd != 0, h.read_len, h.frag_len, h.mate_status == 3);
  qa.matePos = h.mate_pos;
  qa.mateIsFwd = h.mate_fwd != 0;
  qa.mateLen = h.mate_len;
  qa.mateStatus = static_cast<MateStatus>(h.mate_status);
  qa.alnScore(h.aln_score);
  qa.chainStatus = rapmap::utils::FragmentChainStatus(static_cast<ChainStatus>(h.chain_status & 15), static_cast<ChainStatus>(h.chain_status >> 4));
  return qa;
}

// MappingOpts (src/RapMapSAMapper.cpp:114-152; Salmon keeps the same names in SalmonOpts) -> rapmap_cuda_opts_t.
template <typename MappingOptsT>
inline rapmap_cuda_opts_t toCudaOpts(const MappingOptsT& mo) {
  rapmap_cuda_opts_t o;
  rapmap_cuda_opts_default(&o);
  o.max_num_hits = mo.maxNumHits;
  o.quasi_coverage = mo.quasiCov;
  o.sensitive = mo.sensitive;
  o.strict_check = mo.strictCheck;
  o.fuzzy = mo.fuzzy;
  o.sel_aln = mo.selAln;
  o.consensus_slack = mo.consensusSlack;
  o.min_score_fraction = mo.minScoreFraction;
  o.match_score = mo.matchScore;
  o.mismatch_penalty = mo.mismatchPenalty;
  o.gap_open_penalty = mo.gapOpenPenalty;
  o.gap_extend_penalty = mo.gapExtendPenalty;
  o.dp_bandwidth = mo.dpBandwidth;
  o.hard_filter = mo.hardFilter;
  o.alignment_policy = static_cast<uint8_t>(mo.ap);
  o.no_orphans = mo.noOrphans;
  o.no_dovetail = mo.noDovetail;
  o.max_mmp_extension = mo.maxMMPExtension;
  o.recover_orphans = mo.recoverOrphans;
  return o;
}

// One per worker thread, like the SACollector / SASearcher / KSW2Aligner set-up it stands in for
// (src/RapMapSAMapper.cpp:385-455).  Errors throw std::runtime_error with rapmap_cuda_last_error() (the reference logs and
// calls std::exit(1); the caller decides here).
class BatchMapper {
 public:
  BatchMapper(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t& opts, uint64_t maxBatch, uint32_t maxReadLen) : maxBatch_(maxBatch) {
    if (rapmap_cuda_mapper_create(idx, &opts, maxBatch, maxReadLen, &m_) != RAPMAP_OK) throw std::runtime_error(rapmap_cuda_last_error());
    hits_.resize(maxBatch * 8 + 1024);
    off_.resize(maxBatch + 1);
  }
  ~BatchMapper() { rapmap_cuda_mapper_free(m_); }
  BatchMapper(const BatchMapper&) = delete;
  BatchMapper& operator=(const BatchMapper&) = delete;

  // Paired-end chunk (a fastx_parser::ReadGroup<ReadPair>, or any range of pairs with .first.seq / .second.seq):
  // jointHits[i] = what the loop body of processReadsPairSA (src/RapMapSAMapper.cpp:461-702) leaves in `jointHits` for
  // pair i; the five HitCounters are added to hctr.
  template <typename ReadGroupT>
  void mapPairs(ReadGroupT& rg, std::vector<std::vector<rapmap::utils::QuasiAlignment>>& jointHits, rapmap::utils::HitCounters& hctr) {
    seq1_.clear(); seq2_.clear();
    off1_.assign(1, 0); off2_.assign(1, 0);
    for (auto& rp : rg) {
      seq1_.insert(seq1_.end(), rp.first.seq.begin(), rp.first.seq.end());
      off1_.push_back(seq1_.size());
      seq2_.insert(seq2_.end(), rp.second.seq.begin(), rp.second.seq.end());
      off2_.push_back(seq2_.size());
    }
    if (seq1_.empty()) seq1_.push_back(0);
    if (seq2_.empty()) seq2_.push_back(0);
    run(seq1_.data(), off1_.data(), seq2_.data(), off2_.data(), off1_.size() - 1, jointHits, hctr);
  }

  // Unmated chunk (fastx_parser::ReadGroup<ReadSeq>): the loop body of processReadsSingleSA (:156-371).
  template <typename ReadGroupT>
  void mapSingles(ReadGroupT& rg, std::vector<std::vector<rapmap::utils::QuasiAlignment>>& hits, rapmap::utils::HitCounters& hctr) {
    seq1_.clear();
    off1_.assign(1, 0);
    for (auto& r : rg) {
      seq1_.insert(seq1_.end(), r.seq.begin(), r.seq.end());
      off1_.push_back(seq1_.size());
    }
    if (seq1_.empty()) seq1_.push_back(0);
    run(seq1_.data(), off1_.data(), nullptr, nullptr, off1_.size() - 1, hits, hctr);
  }

 private:
  void run(const uint8_t* s1, const uint64_t* o1, const uint8_t* s2, const uint64_t* o2, uint64_t n,
           std::vector<std::vector<rapmap::utils::QuasiAlignment>>& out, rapmap::utils::HitCounters& hctr) {
    if (n > maxBatch_) throw std::runtime_error("chunk larger than the BatchMapper's maxBatch");
    rapmap_read_batch_t rb{s1, o1, s2, o2, n, 0, RAPMAP_LOC_HOST};
    for (;;) {
      rapmap_hit_batch_t hb{hits_.data(), hits_.size(), off_.data(), 0, {0, 0, 0, 0, 0}, RAPMAP_LOC_HOST};
      const int rc = rapmap_cuda_map_batch(m_, &rb, &hb);
      if (rc == RAPMAP_ERR_CAPACITY) { hits_.resize(hb.num_hits + 1024); continue; }
      if (rc != RAPMAP_OK) throw std::runtime_error(rapmap_cuda_last_error());
      hctr.numReads += hb.counters[0]; hctr.peHits += hb.counters[1]; hctr.seHits += hb.counters[2];
      hctr.totHits += hb.counters[3]; hctr.tooManyHits += hb.counters[4];
      break;
    }
    out.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
      auto& v = out[i];
      v.clear();
      for (uint64_t h = off_[i]; h < off_[i + 1]; ++h) v.push_back(toQuasiAlignment(hits_[h]));
    }
  }

  rapmap_cuda_mapper_t* m_{nullptr};
  uint64_t maxBatch_;
  std::vector<rapmap_hit_t> hits_;
  std::vector<uint64_t> off_;
  std::vector<uint8_t> seq1_, seq2_;
  std::vector<uint64_t> off1_, off2_;
};

} // namespace rapmap_b200

#endif // RAPMAP_B200_ADAPTER_HPP
