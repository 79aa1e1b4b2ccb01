// Kernel 3 — mate merging and the per-pair tail of processReadsPairSA, one pair per thread.
//
// Replaces rapmap::utils::mergeLeftRightHits (reference include/RapMapUtils.hpp:1185-1264) and the
// post-merge steps of src/RapMapSAMapper.cpp:533-551,:684-698,:702 (maxNumHits clear, --noOrphans,
// --noDovetail, totHits).  Run twice with the same deterministic logic: pass 0 counts the surviving
// hits per pair (-> exclusive scan -> pair_offsets), pass 1 writes rapmap_hit_t records at their
// final, input-ordered position.  The lists are tiny (mean ~4 per read), so a thread per pair is the
// right grain; both passes stream coalesced QASummary records.
#pragma once
#include "kernels.cuh"

namespace rapmap_b200 {

struct MergeParams {
  DevOpts opts;
  uint64_t numPairs;
  uint8_t pairedInput;
  const QASummary* qsumm;   // reads [0,n) mate 1, [n,2n) mate 2
  const QARec* qaArena;
  const ReadSummary* summ;  // readLen, found
  uint32_t* pairCount;      // pass 0 out
  const uint64_t* pairOffset;  // pass 1 in (exclusive scan of pairCount)
  rapmap_hit_t* hits;       // pass 1 out
  uint64_t hitsCap;
  Counters5* counters;      // pass 0 only
  const int32_t* posPool;   // allPositions / oppositeStrandPositions (fuzzy merge)
};

__device__ __forceinline__ bool dovetailDrop(const rapmap_hit_t& h) {  // src/RapMapSAMapper.cpp:687-696
  if (h.fwd != h.mate_fwd) {
    if (h.fwd && (h.pos > h.mate_pos)) return true;
    if (h.mate_fwd && (h.mate_pos > h.pos)) return true;
  }
  return false;
}

template <bool WRITE>
__device__ __forceinline__ uint32_t mergeSimple(const MergeParams& P, uint64_t pi, unsigned long long* ctr) {
  const DevOpts& o = P.opts;
  QASummary ls = P.qsumm[pi];
  const QARec* L = P.qaArena + ls.qaOff;
  const uint16_t lLen = P.summ[pi].readLen;
  rapmap_hit_t* out = nullptr;
  uint64_t room = 0;
  if (WRITE) {
    uint64_t off = P.pairOffset[pi];
    room = P.pairOffset[pi + 1] - off;
    out = P.hits + off;
    if (off + room > P.hitsCap) return 0;
  }
  uint32_t nOut = 0;
  auto emit = [&](const rapmap_hit_t& h) {
    if (WRITE) { if (nOut < room) out[nOut] = h; }
    ++nOut;
  };
  if (!P.pairedInput) {  // unmated reads: src/RapMapSAMapper.cpp:203-212 (clear when > maxNumHits)
    uint32_t n = ls.nQA;
    if (n > o.maxNumHits) n = 0;
    for (uint32_t i = 0; i < n; ++i) {
      rapmap_hit_t h;
      h.tid = L[i].tid; h.pos = L[i].pos; h.mate_pos = 0; h.frag_len = 0; h.read_len = lLen; h.mate_len = 0; h.aln_score = 0;
      h.fwd = L[i].fwd; h.mate_fwd = 1; h.mate_status = 0; h.chain_status = static_cast<uint8_t>(L[i].chain | (4 << 4));
      emit(h);
    }
    if (!WRITE) { ctr[0] += 1; ctr[3] += ls.nQA; }  // totHits is taken before the clear (:241-246)
    return nOut;
  }
  QASummary rs = P.qsumm[P.numPairs + pi];
  const QARec* R = P.qaArena + rs.qaOff;
  const uint16_t rLen = P.summ[P.numPairs + pi].readLen;
  const uint32_t nl = ls.nQA, nr = rs.nQA;
  bool tooManyHits = false;
  // pass A: count the concordant pairs (two-pointer intersection on tid)
  uint32_t numHits = 0;
  if (nl > 0 && nr > 0) {
    uint32_t li = 0, ri = 0;
    while (li < nl && ri < nr) {
      uint32_t lt = L[li].tid, rt = R[ri].tid;
      if (lt < rt) { ++li; }
      else {
        if (!(rt < lt)) {
          ++numHits;
          if (numHits > o.maxNumHits) { tooManyHits = true; break; }
          ++li;
        }
        ++ri;
      }
    }
  }
  uint32_t nJoint = tooManyHits ? 0 : numHits;
  bool paired = nJoint > 0;
  bool orphans = (!paired) && (nl + nr > 0) && !tooManyHits;
  if (!WRITE) {
    ctr[0] += 1;
    if (tooManyHits) ctr[4] += 1;
    if (paired) ctr[1] += nJoint;
    else if (orphans) ctr[2] += nl + nr;
  }
  uint32_t size = paired ? nJoint : (orphans ? nl + nr : 0);
  if (size > o.maxNumHits) size = 0;                 // src/RapMapSAMapper.cpp:533-536
  if (size > 0 && o.noOrphans && !paired) size = 0;  // :539-551
  if (size == 0) { return 0; }
  if (paired) {
    uint32_t li = 0, ri = 0;
    while (li < nl && ri < nr) {
      uint32_t lt = L[li].tid, rt = R[ri].tid;
      if (lt < rt) { ++li; }
      else {
        if (!(rt < lt)) {
          int32_t s1 = L[li].pos > 0 ? L[li].pos : 0;
          int32_t s2 = R[ri].pos > 0 ? R[ri].pos : 0;
          bool read1First = s1 < s2;
          int32_t fragStart = read1First ? s1 : s2;
          int32_t fragEnd = read1First ? (s2 + static_cast<int32_t>(rLen)) : (s1 + static_cast<int32_t>(lLen));
          rapmap_hit_t h;
          h.tid = lt; h.pos = s1; h.mate_pos = s2; h.frag_len = static_cast<uint32_t>(fragEnd - fragStart);
          h.read_len = lLen; h.mate_len = rLen; h.aln_score = 0; h.fwd = L[li].fwd; h.mate_fwd = R[ri].fwd; h.mate_status = 3;
          h.chain_status = static_cast<uint8_t>(L[li].chain | (R[ri].chain << 4));
          if (!(o.noDovetail && dovetailDrop(h))) emit(h);
          ++li;
        }
        ++ri;
      }
    }
  } else {
    for (int side = 0; side < 2; ++side) {
      const QARec* Q = side ? R : L;
      uint32_t n = side ? nr : nl;
      for (uint32_t i = 0; i < n; ++i) {
        rapmap_hit_t h;
        h.tid = Q[i].tid; h.pos = Q[i].pos; h.mate_pos = 0; h.frag_len = 0; h.read_len = side ? rLen : lLen; h.mate_len = 0; h.aln_score = 0;
        h.fwd = Q[i].fwd; h.mate_fwd = 1; h.mate_status = side ? 2 : 1;
        h.chain_status = side ? static_cast<uint8_t>(4 | (Q[i].chain << 4)) : static_cast<uint8_t>(Q[i].chain | (4 << 4));
        // --noDovetail without selAln (src/RapMapSAMapper.cpp:684-698) reads matePos of orphans, which the reference
        // leaves uninitialised (undefined behaviour); oracle and device both take matePos == 0 there.
        if (!(o.noDovetail && dovetailDrop(h))) emit(h);
      }
    }
  }
  if (!WRITE) ctr[3] += nOut;
  return nOut;
}

template <bool FUZZY>
__global__ void __launch_bounds__(256) merge_count_kernel(MergeParams P) {
  unsigned long long ctr[5] = {0, 0, 0, 0, 0};
  for (uint64_t pi = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; pi < P.numPairs; pi += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    P.pairCount[pi] = mergeSimple<false>(P, pi, ctr);
  }
  // block reduction of the five HitCounters, one atomic per warp
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    unsigned long long v = ctr[c];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.counters->v[c], v);
  }
}

template <bool FUZZY>
__global__ void __launch_bounds__(256) merge_write_kernel(MergeParams P) {
  for (uint64_t pi = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; pi < P.numPairs; pi += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    mergeSimple<true>(P, pi, nullptr);
  }
}

} // namespace rapmap_b200
