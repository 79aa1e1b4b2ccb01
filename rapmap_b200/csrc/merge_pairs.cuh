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
#include "orphan_recovery.cuh"

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
  // orphan recovery (--recoverOrphans)
  BatchView reads;
  const uint8_t* text;
  const int32_t* txpOffsets;
  const int32_t* txpLens;
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

// findBestHitFWRC (include/RapMapUtils.hpp:903-975): closest "rc read downstream of fwd read" pair of positions.
__device__ __forceinline__ bool bestFwRc(const int32_t* fw, uint32_t nFw, const int32_t* rc, uint32_t nRc, int32_t fwdReadLen, int32_t& fwPos,
                                         int32_t& rcPos, int32_t& gapOut) {
  if (nFw == 0 || nRc == 0) return false;
  const int32_t maxGap = 0x7fffffff;
  int32_t bestGap = maxGap;
  uint32_t bf = 0, br = 0;
  for (uint32_t fi = 0; fi < nFw; ++fi) {
    const int32_t p1 = fw[fi];
    uint32_t lo = 0, hi = nRc;  // std::lower_bound
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (rc[mid] < p1) lo = mid + 1; else hi = mid; }
    uint32_t cand[2];
    int nc = 0;
    if (lo == nRc) cand[nc++] = lo - 1;
    else if (lo == 0) cand[nc++] = lo;
    else { cand[nc++] = lo; cand[nc++] = lo - 1; }
    for (int c = 0; c < nc; ++c) {
      const int32_t r = rc[cand[c]];
      int32_t d = r - (p1 + fwdReadLen);
      int32_t gap = (r >= p1) ? (d < 0 ? -d : d) : maxGap;
      if (gap < bestGap) { bestGap = gap; bf = fi; br = cand[c]; }
    }
  }
  if (bestGap == maxGap) return false;
  fwPos = fw[bf]; rcPos = rc[br]; gapOut = bestGap;
  return true;
}

// recoverSingleOrphan of selective_alignment::utils::recoverOrphans (include/SelectiveAlignmentUtils.hpp:69-176): looks for
// the other read end in the <= 1000 bases downstream (anchor forward: the reverse complement of the other end) or
// upstream (anchor reverse: the other end as it is) of the anchor hit.  Plain arguments only: the function is out of line.
__device__ __noinline__ bool recoverOne(const uint8_t* text, const int32_t* txpOffsets, const int32_t* txpLens, const uint8_t* otherRead, int32_t otherLen,
                                        int32_t anchorLen, int32_t anchorPos, bool anchorFwd, uint32_t tid, int32_t& otherPos) {
  const int64_t toff = txpOffsets[tid];
  const int32_t refLength = txpLens[tid];
  int32_t startPos, windowLength;
  if (anchorFwd) {
    startPos = anchorPos > 0 ? anchorPos : 0;
    windowLength = min(1000, refLength - startPos);
  } else {
    const int32_t endPos = min(refLength, anchorPos + anchorLen);
    startPos = endPos - 1000 > 0 ? endPos - 1000 : 0;
    windowLength = min(1000, endPos);
  }
  int firstEnd;
  const int d = semiGlobalMyers(otherRead, otherLen, anchorFwd, text + toff + startPos, windowLength, otherLen / 4, firstEnd);
  if (d < 0) return false;
  otherPos = startPos + firstEnd - otherLen;
  return true;
}

// mergeLeftRightHitsFuzzy (include/RapMapUtils.hpp:864-1183) + the post-merge steps of src/RapMapSAMapper.cpp:533-551.
// With selAln the hits are scored and filtered afterwards (sel_aln.cuh); --noDovetail is applied there.
template <bool WRITE>
__device__ __forceinline__ uint32_t mergeFuzzy(const MergeParams& P, uint64_t pi, unsigned long long* ctr) {
  const DevOpts& o = P.opts;
  const QASummary ls = P.qsumm[pi], rs = P.qsumm[P.numPairs + pi];
  const QARec* L = P.qaArena + ls.qaOff;
  const QARec* R = P.qaArena + rs.qaOff;
  const ReadSummary lsum = P.summ[pi], rsum = P.summ[P.numPairs + pi];
  const uint16_t lLen = lsum.readLen, rLen = rsum.readLen;
  const uint32_t nl = ls.nQA, nr = rs.nQA;
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
  auto orphan = [&](const QARec& q, int side) {
    rapmap_hit_t h;
    h.tid = q.tid; h.pos = q.pos; h.mate_pos = 0; h.frag_len = 0; h.read_len = side ? rLen : lLen; h.mate_len = 0; h.aln_score = 0;
    h.fwd = q.fwd; h.mate_fwd = 1; h.mate_status = side ? 2 : 1;
    h.chain_status = side ? static_cast<uint8_t>(4 | (q.chain << 4)) : static_cast<uint8_t>(q.chain | (4 << 4));
    return h;
  };
  // --recoverOrphans (src/RapMapSAMapper.cpp:498-530): the orphans of the listed sides are not reported; every one of them
  // is an anchor next to which the other read end is searched, and the pairs found this way are the pair's hits.
  auto runRecovery = [&](uint32_t nlR, uint32_t nrR) -> uint32_t {
    if (WRITE && room == 0) return 0;
    const uint8_t* reads[2];
    for (int e = 0; e < 2; ++e) {
      const uint64_t ri = pi;
      reads[e] = P.reads.off[e] ? P.reads.seq[e] + P.reads.off[e][ri] : P.reads.seq[e] + ri * P.reads.fixedLen;
    }
    uint32_t nRec = 0;
    auto rec = [&](const QARec& a, bool anchorIsLeft) {
      const int32_t anchorPos = P.posPool[a.posOff];  // allPositions.front()
      const int32_t otherLen = anchorIsLeft ? rLen : lLen, anchorLen = anchorIsLeft ? lLen : rLen;
      int32_t otherPos = -1;
      if (!recoverOne(P.text, P.txpOffsets, P.txpLens, reads[anchorIsLeft ? 1 : 0], otherLen, anchorLen, anchorPos, a.fwd != 0, a.tid, otherPos)) return;
      ++nRec;
      const int32_t lpos = anchorIsLeft ? anchorPos : otherPos, rpos = anchorIsLeft ? otherPos : anchorPos;
      const int32_t s1 = lpos > 0 ? lpos : 0, s2 = rpos > 0 ? rpos : 0;
      const bool read1First = s1 < s2;
      const int32_t fragStart = read1First ? s1 : s2;
      const int32_t fragEnd = read1First ? (s2 + static_cast<int32_t>(rLen)) : (s1 + static_cast<int32_t>(lLen));
      rapmap_hit_t h;
      h.tid = a.tid; h.pos = lpos; h.mate_pos = rpos; h.frag_len = static_cast<uint32_t>(fragEnd - fragStart);
      h.read_len = lLen; h.mate_len = static_cast<uint16_t>(otherLen);  // the reference stores the OTHER end's length as mateLen
      h.aln_score = 0;
      h.fwd = anchorIsLeft ? a.fwd : !a.fwd; h.mate_fwd = anchorIsLeft ? !a.fwd : a.fwd; h.mate_status = 3;
      h.chain_status = anchorIsLeft ? static_cast<uint8_t>(a.chain | (4 << 4)) : static_cast<uint8_t>(4 | (a.chain << 4));
      if (!(!o.selAln && o.noDovetail && dovetailDrop(h))) emit(h);
    };
    uint32_t li = 0, ri = 0;
    while (li < nlR && ri < nrR) {
      const uint32_t lt = L[li].tid, rt = R[ri].tid;
      if (lt < rt) rec(L[li++], true);
      else if (rt < lt) rec(R[ri++], false);
      else { ++li; ++ri; }  // a shared transcript cannot occur here (the reference exits on it)
    }
    while (li < nlR) rec(L[li++], true);
    while (ri < nrR) rec(R[ri++], false);
    if (nRec > o.maxNumHits) return 0;  // src/RapMapSAMapper.cpp:533-536
    if (!WRITE && !o.selAln) ctr[3] += nOut;
    return nOut;
  };
  bool tooManyHits = false;
  int mode = 0;  // 0 nothing, 1 only right, 2 only left, 3 paired
  uint32_t numHits = 0, sameTxp = 0;
  if (nl == 0) {
    if (!lsum.found && nr > 0) mode = 1;   // orphans only if the other end had no k-mer hit at all (:880-899)
  } else if (nr == 0) {
    if (!rsum.found) mode = 2;
  } else {
    mode = 3;
  }
  if (!WRITE) ctr[0] += 1;
  if (mode == 1 || mode == 2) {
    const uint32_t n = mode == 1 ? nr : nl;
    if (!WRITE) { ctr[2] += n; ctr[1] += n; }  // seHits, and peHits counts every non-empty jointHits (:1176-1179)
    if (o.recoverOrphans) return runRecovery(mode == 2 ? nl : 0u, mode == 1 ? nr : 0u);
    uint32_t size = n;
    if (size > o.maxNumHits) size = 0;
    if (o.noOrphans) size = 0;
    if (size == 0) return 0;
    const QARec* Q = mode == 1 ? R : L;
    for (uint32_t i = 0; i < n; ++i) {
      rapmap_hit_t h = orphan(Q[i], mode == 1 ? 1 : 0);
      if (!o.selAln && o.noDovetail && dovetailDrop(h)) continue;
      emit(h);
    }
    if (!WRITE && !o.selAln) ctr[3] += nOut;
    return nOut;
  }
  if (mode == 0) return 0;
  // paired: the walk is run twice in count mode (once to learn tooManyHits), once more to write
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t li = 0, ri = 0;
    numHits = 0;
    const bool writing = pass == 1;
    while (li < nl && ri < nr) {
      const uint32_t lt = L[li].tid, rt = R[ri].tid;
      if (lt < rt) { ++li; }
      else {
        if (!(rt < lt)) {
          ++sameTxp;
          const QARec& l = L[li];
          const QARec& r = R[ri];
          const int32_t* lAll = P.posPool + l.posOff; const int32_t* lOpp = P.posPool + l.oppOff;
          const int32_t* rAll = P.posPool + r.posOff; const int32_t* rOpp = P.posPool + r.oppOff;
          const int32_t* leftFwd = l.fwd ? lAll : lOpp;   const uint32_t nLeftFwd = l.fwd ? l.nAll : l.nOpp;
          const int32_t* leftRC = l.fwd ? lOpp : lAll;    const uint32_t nLeftRC = l.fwd ? l.nOpp : l.nAll;
          const int32_t* rightFwd = r.fwd ? rAll : rOpp;  const uint32_t nRightFwd = r.fwd ? r.nAll : r.nOpp;
          const int32_t* rightRC = r.fwd ? rOpp : rAll;   const uint32_t nRightRC = r.fwd ? r.nOpp : r.nAll;
          int32_t a1 = 0, a2 = 0, ag = 0, b1 = 0, b2 = 0, bg = 0;
          const bool haveFWRC = bestFwRc(leftFwd, nLeftFwd, rightRC, nRightRC, static_cast<int32_t>(lLen), a1, a2, ag);
          const bool haveRCFW = bestFwRc(rightFwd, nRightFwd, leftRC, nLeftRC, static_cast<int32_t>(rLen), b1, b2, bg);
          bool found = false, leftFwdFlag = false, rightFwdFlag = false;
          int32_t bestGap = 0x7fffffff, leftPos = -1, rightPos = -1;
          if (haveFWRC) { leftPos = a1; rightPos = a2; bestGap = ag; leftFwdFlag = true; rightFwdFlag = false; found = true; }
          if (haveRCFW) {
            if (bg < bestGap) { leftPos = b2; rightPos = b1; leftFwdFlag = false; rightFwdFlag = true; }
            found = true;
          }
          if (found) {
            ++numHits;
            if (writing) {
              const int32_t s1 = leftPos > 0 ? leftPos : 0, s2 = rightPos > 0 ? rightPos : 0;
              const bool read1First = s1 < s2;
              const int32_t fragStart = read1First ? s1 : s2;
              const int32_t fragEnd = read1First ? (s2 + static_cast<int32_t>(rLen)) : (s1 + static_cast<int32_t>(lLen));
              rapmap_hit_t h;
              h.tid = lt; h.pos = leftPos; h.mate_pos = rightPos; h.frag_len = static_cast<uint32_t>(fragEnd - fragStart);
              h.read_len = lLen; h.mate_len = rLen; h.aln_score = 0; h.fwd = leftFwdFlag; h.mate_fwd = rightFwdFlag; h.mate_status = 3;
              h.chain_status = static_cast<uint8_t>(l.chain | (r.chain << 4));
              if (!(!o.selAln && o.noDovetail && dovetailDrop(h))) emit(h);
            }
            if (numHits > o.maxNumHits) { tooManyHits = true; break; }
          }
          ++li;
        }
        ++ri;
      }
    }
    if (pass == 0) {
      uint32_t nJoint = tooManyHits ? 0 : numHits;
      if (!WRITE) { if (tooManyHits) ctr[4] += 1; ctr[1] += nJoint; }
      if (o.recoverOrphans && !tooManyHits && numHits == 0 && sameTxp == 0) return runRecovery(nl, nr);  // HAD_EMPTY_INTERSECTION
      if (nJoint == 0 || nJoint > o.maxNumHits) return 0;
    }
  }
  if (!WRITE && !o.selAln) ctr[3] += nOut;
  return nOut;
}

template <bool FUZZY>
__global__ void __launch_bounds__(256) merge_count_kernel(MergeParams P) {
  unsigned long long ctr[5] = {0, 0, 0, 0, 0};
  for (uint64_t pi = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; pi < P.numPairs; pi += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    P.pairCount[pi] = FUZZY ? mergeFuzzy<false>(P, pi, ctr) : mergeSimple<false>(P, pi, ctr);
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
    if (FUZZY) mergeFuzzy<true>(P, pi, nullptr); else mergeSimple<true>(P, pi, nullptr);
  }
}

} // namespace rapmap_b200
