// Kernel 1 — "the SA-lookup kernel": seed + maximal-mappable-prefix collection for one read per warp.
//
// Replaces SACollector::operator() / getSAHits_ / spotCheck_ (reference include/SACollector.hpp:108-362,
// :441-677, :366-431) and SASearcher::extendSearchNaive (include/SASearcher.hpp:87-309) for the default
// strand-decision mode (disableNIP && strictCheck => coverage check, SACollector.hpp:138,:283-288), with
// or without chain scoring (selAln).
//
// B200 mapping: one warp owns one read.  The reference's walk is a chain of dependent random reads
// (hash probe -> SA probe -> text compare).  The walk itself must be replayed decision by decision to
// stay bit-exact, so the warp shortens the chain instead:
//   * k-mer lookups are pure, so the 32 lanes look up 16 consecutive read positions in BOTH strands at
//     once (one DRAM round trip instead of up to 31 sequential ones across a mismatch) and park the
//     intervals in shared memory; the rc-strand walk reuses the same entries (k-mer at p of the reverse
//     complement == RC of the k-mer at L-k-p).
//   * every suffix comparison of the binary searches covers 128 text bytes per round trip (4 per lane),
//     reduced with one ballot.
// All walk state is replicated in registers across the warp (warp-uniform control flow).
#pragma once
#include "kernels.cuh"

namespace rapmap_b200 {

struct CollectParams {
  DeviceIndex ix;
  BatchView reads;
  DevOpts opts;
  uint32_t maxReadLen;     // shared-memory sizing
  uint32_t lpad;           // bytes per read buffer (multiple of 16)
  uint32_t pmax;           // max k-mer positions per read
  uint32_t warpSmemBytes;
  uint32_t packOff;        // offset of the packed-read area inside a warp's shared-memory slice
  uint32_t ctxOff;         // offset of the per-warp WarpCtx
  uint32_t voteOff;        // offset of the per-position k-mer vote arrays (strand decision without coverage)
  ReadSummary* summ;
  IntervalRec* arena;
  uint32_t arenaCap;
  uint32_t* arenaCursor;
  uint32_t* status;
};

__device__ __forceinline__ int baseCode(uint8_t c) {  // reference include/Kmer.hpp:40-51
  switch (c | 0x20) {
    case 'a': return 0;
    case 'c': return 1;
    case 'g': return 2;
    case 't': return 3;
    default: return -1;
  }
}

// Kmer::fromChars (include/Kmer.hpp:524-542): stops at the first invalid base, leaving the partial word.
__device__ __noinline__ bool encodeKmer(const uint8_t* s, int k, uint64_t& w) {
  w = 0;
  int shift = 2 * k - 2;
  for (int i = 0; i < k; ++i, shift -= 2) {
    int c = baseCode(s[i]);
    if (c < 0) return false;
    w |= static_cast<uint64_t>(c) << shift;
  }
  return true;
}

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

struct WarpCtx {
  DeviceIndex ix;
  const uint8_t* fwdBuf;
  const uint8_t* rcBuf;
  int2* cF;   // per forward position: interval of the k-mer        (x == -2: not looked up, x == -1: absent)
  int2* cR;   // per forward position: interval of its reverse complement
  const uint64_t* packF;  // 2-bit packed read, base i at bits (63-2(i%32), 62-2(i%32)) of word i/32 (forward strand)
  const uint64_t* packR;  // same for the reverse complement
  const uint32_t* invF;   // bit i%32 of word i/32: base i is not A/C/G/T (either case)
  const uint32_t* invR;
  const uint32_t* nmF;    // bit set: base is N/n
  const uint32_t* nmR;
  int L, k, npos, lane;
  int maxMMPExtension, maxInterval;
  int8_t* voteF;   // per forward position: +1 present / -1 absent / 0 untested, forward orientation (KmerDirScore::fwdScore)
  int8_t* voteR;   // same for the reverse-complement orientation
  bool doChaining;
  bool hasU;
  bool disableNIP, voteMode;
};

// k-mer at position p of a packed strand in O(1): funnel shift of two packed words; valid == no non-ACGT base in
// the window.  Invalid windows fall back to encodeKmer (partial-word semantics of Kmer::fromChars).
__device__ __forceinline__ bool packedKmer(const uint64_t* pack, const uint32_t* inv, const uint8_t* buf, int p, int k, uint64_t& w) {
  const int wi = p >> 5, sh = p & 31;
  const uint32_t m0 = inv[wi], m1 = inv[wi + 1];
  const uint32_t win = sh ? ((m0 >> sh) | (m1 << (32 - sh))) : m0;
  if ((win & ((1u << k) - 1u)) != 0u) return encodeKmer(buf + p, k, w);
  const uint64_t w0 = pack[wi], w1 = pack[wi + 1];
  const uint64_t val = sh ? ((w0 << (2 * sh)) | (w1 >> (64 - 2 * sh))) : w0;
  w = val >> (64 - 2 * k);
  return true;
}

__device__ __forceinline__ int findNMask(const uint32_t* nm, int from, int L) {  // std::string::find_first_of("nN", from)
  if (from >= L) return 0x7fffffff;
  int wi = from >> 5;
  uint32_t m = nm[wi] & (0xffffffffu << (from & 31));
  while (true) {
    if (m) return wi * 32 + __ffs(m) - 1;
    ++wi;
    if (wi * 32 >= L) return 0x7fffffff;
    m = nm[wi];
  }
}

// Speculative lookup of 16 consecutive forward positions q0, q0+dir, ... in both orientations.
#ifndef RAPMAP_FILL_ATTR
#define RAPMAP_FILL_ATTR __forceinline__
#endif
__device__ RAPMAP_FILL_ATTR void fillCache(const WarpCtx& c, int q0, int dir) {
  int q = q0 + dir * ((static_cast<int>(threadIdx.x) & 31) & 15);
  bool rcSide = ((static_cast<int>(threadIdx.x) & 31) >> 4) != 0;
  if (q >= 0 && q < c.npos) {
    int2* slot = (rcSide ? c.cR : c.cF) + q;
    if (slot->x == -2) {
      uint64_t w;
      int2 res = make_int2(-1, -1);
      if (packedKmer(c.packF, c.invF, c.fwdBuf, q, c.k, w)) res = hashFind(c.ix, rcSide ? kmerRC(w, c.k) : w);
      *slot = res;
    }
  }
  __syncwarp();
}

// Lookup of the k-mer `w` found at position p of the walk strand (and of its reverse complement).
// Full, cacheable k-mers go through the shared-memory cache; partial words (a non-ACGT base inside the
// window, Kmer.hpp:536-537) and reads containing 'U' (reverseRead maps U->A, so strand symmetry breaks)
// are looked up directly.
__device__ __forceinline__ void lookupBoth(const WarpCtx& c, bool isRC, int p, uint64_t w, bool valid, bool needMer, bool needComp,
                                           int2& mer, int2& comp) {
  if (valid && !(isRC && c.hasU)) {
    int q = isRC ? (c.L - c.k - p) : p;
    int2* cm = isRC ? c.cR : c.cF;
    int2* cc = isRC ? c.cF : c.cR;
    if ((needMer && cm[q].x == -2) || (needComp && cc[q].x == -2)) fillCache(c, q, isRC ? -1 : 1);
    mer = cm[q];
    comp = cc[q];
  } else {
    if (needMer) mer = hashFind(c.ix, w);
    if (needComp) comp = hashFind(c.ix, kmerRC(w, c.k));
  }
}

// Cooperative suffix comparison: query q[i] vs text[t+i] for i >= i0 while i < m and t+i < n.
// sentIdx >= 0 replaces q[sentIdx] by `sent` (searches 2/3 of extendSearchNaive).  Returns the index at
// which the reference's inner while-loop stops; rel = -1 (query < text), +1 (query > text), 0 (ran off).
__device__ __noinline__ int coopCompare(const WarpCtx& c, const uint8_t* q, int m, int64_t t, int i0, int sentIdx, uint8_t sent, int& rel) {
  int64_t limL = c.ix.n - t;
  int lim = (limL < static_cast<int64_t>(m)) ? static_cast<int>(limL) : m;
  if (i0 >= lim) { rel = 0; return i0; }
  const uint8_t* tp = c.ix.text + t;
  for (int base = i0;; base += 128) {
    int idx = base + (static_cast<int>(threadIdx.x) & 31) * 4;
    uint8_t tc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc[j] = (idx + j < lim) ? __ldg(tp + idx + j) : 0;
    int stop = 4, r = 0;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      int id = idx + j;
      if (id >= lim) { stop = j; r = 0; }
      else {
        uint8_t qc = (id == sentIdx) ? sent : upperChar(q[id]);
        if (qc != tc[j]) { stop = j; r = (qc < tc[j]) ? -1 : 1; }
      }
    }
    unsigned b = __ballot_sync(0xffffffffu, stop < 4);
    if (b) {
      int src = __ffs(b) - 1;
      int sIdx = __shfl_sync(0xffffffffu, idx + stop, src);
      rel = __shfl_sync(0xffffffffu, r, src);
      return sIdx;
    }
  }
}

// Suffix table.  For an SA range of 2..32 suffixes, lane i takes suffix i: one coalesced SA load, then its own common
// prefix with the query, 8 bytes per step.  (lcp, order relation at the first difference, SA value) stay in that lane's
// registers and the three binary searches of extendSearchNaive are replayed with shuffles instead of memory probes:
// ~0.4x the instructions of probing suffix by suffix with the whole warp.  (A first version that let all 32 lanes
// cooperate on every suffix was 1.6x SLOWER than no table at all: this kernel is bound by instruction issue and
// instruction fetch, not by memory latency - profiles/r01_notes.md.)
struct SufTab {
  int64_t lo;   // SA index of the suffix held by lane 0
  int cnt;      // 0 = no table
  int mTab;     // query length the table was built for
  int32_t sa;   // per lane
  int ell;      // per lane: first index >= k where query and suffix differ, capped at min(mTab, n - sa)
  int rel;      // per lane: -1 query < text, +1 query > text, 0 ran off
};

__device__ __forceinline__ uint64_t load8Global(const uint8_t* p) {
  const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(7));
  const unsigned sh = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p) & 7) * 8;
  const uint64_t lo = __ldg(a), hi = __ldg(a + 1);
  return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}
__device__ __forceinline__ uint64_t load8Shared(const uint8_t* p) {
  const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(7));
  const unsigned sh = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p) & 7) * 8;
  const uint64_t lo = a[0], hi = a[1];
  return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}

// q must point into an UPPER-CASED read buffer in shared memory.
__device__ __noinline__ SufTab buildSufTab(const WarpCtx& c, int64_t lbIn, int64_t ubIn, const uint8_t* q, int mTab) {
  SufTab tb;
  tb.lo = lbIn + 1; tb.cnt = 0; tb.mTab = mTab; tb.sa = 0; tb.ell = 0; tb.rel = 0;
  const int64_t cnt64 = ubIn - lbIn - 1;
#ifdef RAPMAP_NO_SUFTAB
  return tb;
#endif
  if (cnt64 < 2 || cnt64 > 32) return tb;
  const int lane = static_cast<int>(threadIdx.x) & 31;
  if (lane < static_cast<int>(cnt64)) {
    const int32_t t = __ldg(c.ix.SA + tb.lo + lane);
    const int64_t limL = c.ix.n - static_cast<int64_t>(t);
    const int lim = limL < static_cast<int64_t>(mTab) ? static_cast<int>(limL) : mTab;
    int ell = c.k, rel = 0;
    const uint8_t* tp = c.ix.text + t;
    while (ell < lim) {
      const uint64_t tw = load8Global(tp + ell), qw = load8Shared(q + ell);
      const uint64_t x = tw ^ qw;
      if (x) {
        const int d = (__ffsll(static_cast<long long>(x)) - 1) >> 3;
        ell += d;
        if (ell < lim) rel = ((qw >> (8 * d)) & 0xffu) < ((tw >> (8 * d)) & 0xffu) ? -1 : 1;
        break;
      }
      ell += 8;
    }
    if (ell >= lim) { ell = lim; rel = 0; }
    tb.sa = t; tb.ell = ell; tb.rel = rel;
  }
  __syncwarp();
  tb.cnt = static_cast<int>(cnt64);
  return tb;
}

// One suffix comparison of extendSearchNaive: from the table when possible, else from memory.
__device__ __forceinline__ int probeSuffix(const WarpCtx& c, const SufTab& tb, int64_t cc, const uint8_t* q, int m, int i0, int sentIdx, uint8_t sent,
                                           int& rel, int64_t& t) {
  if (tb.cnt > 0 && m <= tb.mTab + 1) {
    const int sidx = static_cast<int>(cc - tb.lo);
    if (sidx >= 0 && sidx < tb.cnt) {
      t = __shfl_sync(0xffffffffu, tb.sa, sidx);
      const int ell = __shfl_sync(0xffffffffu, tb.ell, sidx);
      const int rl = __shfl_sync(0xffffffffu, tb.rel, sidx);
      const int64_t limL = c.ix.n - t;
      const int lim = limL < static_cast<int64_t>(m) ? static_cast<int>(limL) : m;
      if (i0 >= lim) { rel = 0; return i0; }
      if (i0 <= ell) {
        if (sentIdx >= 0 && ell >= sentIdx) {
          if (sentIdx < lim) { rel = sent == '#' ? -1 : 1; return sentIdx; }  // '#' < '$' < letters < '{'
          rel = 0;
          return lim;
        }
        if (ell >= lim) { rel = 0; return lim; }
        rel = rl;
        return ell;
      }
    }
  }
  t = __ldg(c.ix.SA + cc);
  return coopCompare(c, q, m, t, i0, sentIdx, sent, rel);
}

// SASearcher::extendSearchNaive (include/SASearcher.hpp:87-309); startAt = k.
__device__ __noinline__ void extendSearch(const WarpCtx& c, const SufTab& tb, int64_t lbIn, int64_t ubIn, const uint8_t* q, int mQ,
                                             int& outLb, int& outUb, int& outLen) {
  const int startAt = c.k;
  const int64_t n = c.ix.n;
  int rel;
  int64_t t;
  if (ubIn - lbIn == 2) {  // :109-126
    int i = probeSuffix(c, tb, lbIn + 1, q, mQ, startAt, -1, 0, rel, t);
    outLb = static_cast<int>(lbIn + 1); outUb = static_cast<int>(ubIn); outLen = i;
    return;
  }
  int64_t l = lbIn, r = ubIn;
  int lcpLP = startAt, lcpRP = startAt;
  int prevILow = startAt, prevIHigh = startAt;
  int maxLen = startAt;
  // Each search halves [l, r]; 80 rounds can only be exceeded on a malformed index (guards the GPU against a hang).
  for (int guard = 0; guard < 80; ++guard) {  // :150-209
    int64_t cc = (l + r) / 2;
    int i0 = lcpLP < lcpRP ? lcpLP : lcpRP;
    int i = probeSuffix(c, tb, cc, q, mQ, i0, -1, 0, rel, t);
    bool plt = true;
    if (rel < 0) { if (i > prevIHigh) prevIHigh = i; }
    else if (rel > 0) { if (i > prevILow) prevILow = i; plt = false; }
    else if (i == mQ || t + i == n) { if (i > prevIHigh) prevIHigh = i; }
    if (plt) {
      if (cc == l + 1) { maxLen = max(max(i, prevILow), prevIHigh); break; }
      r = cc; lcpRP = i;
    } else {
      if (cc == r - 1) { maxLen = max(max(i, prevILow), prevIHigh); break; }
      l = cc; lcpLP = i;
    }
  }
  const int m = maxLen + 1;  // :212
  int64_t bound[2];
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {  // :224-258 lower bound with '#', :270-304 upper bound with '{'
    const uint8_t sent = pass == 0 ? '#' : '{';
    l = pass == 0 ? lbIn : bound[0] - 1;
    r = ubIn;
    lcpLP = startAt; lcpRP = startAt;
    bound[pass] = r;
    for (int guard = 0; guard < 80; ++guard) {
      int64_t cc = (l + r) / 2;
      int i0 = lcpLP < lcpRP ? lcpLP : lcpRP;
      int i = probeSuffix(c, tb, cc, q, m, i0, m - 1, sent, rel, t);
      if (rel <= 0) {
        if (cc == l + 1) { bound[pass] = cc; break; }
        r = cc; lcpRP = i;
      } else {
        if (cc == r - 1) { bound[pass] = r; break; }
        l = cc; lcpLP = i;
      }
    }
  }
  if (bound[0] == bound[1]) bound[1] += 1;  // :307
  outLb = static_cast<int>(bound[0]); outUb = static_cast<int>(bound[1]); outLen = maxLen;
}

__device__ __forceinline__ int findN(const uint8_t* s, int from, int L) {  // std::string::find_first_of("nN", from)
  for (int i = from; i < L; ++i)
    if ((s[i] | 0x20) == 'n') return i;
  return 0x7fffffff;
}

// kmerScores.emplace_back of spotCheck_ / the first-hit scan (include/SACollector.hpp:200-227,:417-430): statuses in
// forward orientation at the forward position; duplicates of a position carry the same statuses, the first one is kept.
__device__ __forceinline__ void recordVote(const WarpCtx& c, bool isRC, int p, bool merPresent, bool compPresent) {
  if (!c.voteMode) return;
  const int q = isRC ? (c.L - c.k - p) : p;
  if ((static_cast<int>(threadIdx.x) & 31) == 0 && c.voteF[q] == 0) {
    c.voteF[q] = (isRC ? compPresent : merPresent) ? 1 : -1;
    c.voteR[q] = (isRC ? merPresent : compPresent) ? 1 : -1;
  }
}

// SASearcher::lce (include/SASearcher.hpp:318-334) including its doubled start offset: the comparison runs at
// SA[p] + startAt + len with len starting at startAt.  32 positions per round trip.
__device__ __noinline__ int lceCoop(const WarpCtx& c, int64_t p1, int64_t p2, int startAt, int stopAt) {
  const int lane = static_cast<int>(threadIdx.x) & 31;
  const int64_t n = c.ix.n;
  p1 = p1 < 0 ? 0 : (p1 >= n ? n - 1 : p1);
  p2 = p2 < 0 ? 0 : (p2 >= n ? n - 1 : p2);
  const int64_t o1 = static_cast<int64_t>(__ldg(c.ix.SA + p1)) + startAt, o2 = static_cast<int64_t>(__ldg(c.ix.SA + p2)) + startAt;
  const int64_t maxIndex = o1 > o2 ? o1 : o2;
  for (int len = startAt;; len += 32) {
    const int idx = len + lane;
    const bool inRange = maxIndex + idx < n;
    const uint8_t a = inRange ? __ldg(c.ix.text + o1 + idx) : 0;
    const uint8_t b = inRange ? __ldg(c.ix.text + o2 + idx) : 1;
    const bool stop = !inRange || a != b || a == '$' || idx >= stopAt;
    const unsigned m = __ballot_sync(0xffffffffu, stop);
    if (m) return len + __ffs(m) - 1;
  }
}

// SACollector::getSAHits_ (include/SACollector.hpp:441-677) on one strand.  `buf` is the strand's read.
__device__ __noinline__ void walkStrand(const WarpCtx& c, bool isRC, int startPos, bool haveStart, int2 startIv,
                                           uint32_t& cov, uint32_t& strandHits, uint32_t& otherStrandHits, IntervalRec* list, int& nList) {
  const uint8_t* buf = isRC ? c.rcBuf : c.fwdBuf;
  const uint64_t* pack = isRC ? c.packR : c.packF;
  const uint32_t* inv = isRC ? c.invR : c.invF;
  const uint32_t* nm = isRC ? c.nmR : c.nmF;
  const int k = c.k, L = c.L;
  int rb = 0;
  int64_t lb = 0, ub = 0;
  bool lastSearch = false;
  int prevMMPEnd = 0;
  bool skipSetup = haveStart;
  if (skipSetup) { rb = startPos; lb = startIv.x; ub = startIv.y; }
  while (skipSetup || rb + k <= L) {
    if (!skipSetup) {
      uint64_t mer;
      bool valid = packedKmer(pack, inv, buf, rb, k, mer);
      if (!valid) {  // :505-516
        int ip = findNMask(nm, rb, L);
        if (ip < rb + k) { rb = ip + 1; continue; }
      }
      if (isHomopolymer(mer, k)) { rb += 1; continue; }  // :520-536
      int2 fm, fc;
      lookupBoth(c, isRC, rb, mer, valid, true, true, fm, fc);  // find + spotCheck_ complement (:541-546,:671)
      if (fm.x >= 0) ++strandHits;
      if (fc.x >= 0) ++otherStrandHits;
      recordVote(c, isRC, rb, fm.x >= 0, fc.x >= 0);
      if (fm.x < 0) { rb += 1; continue; }  // :673
      lb = fm.x; ub = fm.y;
    }
    skipSetup = false;
    lb = lb - 1 > 0 ? lb - 1 : 0;  // :553
    bool firstAttempt = c.doChaining ? (rb == 0) : true;
    int endPos = firstAttempt ? L : min(rb + k + c.maxMMPExtension, L);
    int nlb, nub, matchedLen;
    const SufTab tb = buildSufTab(c, lb, ub, buf + rb, endPos - rb);
    extendSearch(c, tb, lb, ub, buf + rb, endPos - rb, nlb, nub, matchedLen);
    if (c.doChaining && firstAttempt && !(matchedLen >= L) && matchedLen >= k + c.maxMMPExtension) {  // :568-575
      endPos = min(rb + k + c.maxMMPExtension, L);
      extendSearch(c, tb, lb, ub, buf + rb, endPos - rb, nlb, nub, matchedLen);
    }
    lb = nlb; ub = nub;
    if (ub > lb && (ub - lb) < c.maxInterval) {  // :578
      if ((static_cast<int>(threadIdx.x) & 31) == 0) {
        IntervalRec rec;
        rec.begin = static_cast<int32_t>(lb); rec.end = static_cast<int32_t>(ub);
        rec.len = static_cast<uint16_t>(matchedLen); rec.qpos = static_cast<uint16_t>(rb);
        list[nList] = rec;
      }
      ++nList;
      int correction = prevMMPEnd > rb ? prevMMPEnd - rb : 0;
      cov += static_cast<uint32_t>(matchedLen - correction);
      prevMMPEnd = rb + matchedLen;
      if (rb + matchedLen < L) {  // :599-616 mismatching k-mer, both orientations
        int kp = rb + matchedLen - (k - 1);
        uint64_t mm;
        if (packedKmer(pack, inv, buf, kp, k, mm)) {
          int2 fm, fc;
          lookupBoth(c, isRC, kp, mm, true, true, true, fm, fc);
          if (fm.x >= 0) ++strandHits;
          if (fc.x >= 0) ++otherStrandHits;
          recordVote(c, isRC, kp, fm.x >= 0, fc.x >= 0);
        }
      }
    }
    if (lastSearch) return;            // :623
    if (rb + matchedLen >= L) return;  // :630
    {  // :634-657 next start: MMP skip, or the NIP skip when --noSensitive
      const int mismatchPos = rb + matchedLen;
      const int lce = c.disableNIP ? matchedLen : lceCoop(c, lb, ub - 1, matchedLen, L - mismatchPos);
      const int skipMatch = mismatchPos - (k - 1), skipLCE = rb + lce - (k - 1);
      rb = skipMatch > skipLCE ? skipMatch : skipLCE;
      if (!c.disableNIP && lce > matchedLen && L > k) rb = rb < L - k ? rb : L - k;
    }
    if (rb + k == L) lastSearch = true;  // :663
  }
}

#ifndef RAPMAP_K1_MIN_BLOCKS
#define RAPMAP_K1_MIN_BLOCKS 4  // 64 registers per thread: 32 resident warps per SM
#endif
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, RAPMAP_K1_MIN_BLOCKS) sa_collect_kernel(CollectParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t* base = smem + static_cast<size_t>(warp) * P.warpSmemBytes;
  uint8_t* fwdBuf = base;
  uint8_t* rcBuf = base + P.lpad;
  int2* cF = reinterpret_cast<int2*>(base + 2 * P.lpad);
  int2* cR = cF + P.pmax;
  IntervalRec* ivF = reinterpret_cast<IntervalRec*>(cR + P.pmax);
  IntervalRec* ivR = ivF + P.pmax;
  const uint32_t pw = P.lpad / 32 + 2;  // packed words / mask words per strand (one spare word for the funnel shift)
  uint64_t* packF = reinterpret_cast<uint64_t*>(base + P.packOff);
  uint64_t* packR = packF + pw;
  uint32_t* invF = reinterpret_cast<uint32_t*>(packR + pw);
  uint32_t* invR = invF + pw;
  uint32_t* nmF = invR + pw;
  uint32_t* nmR = nmF + pw;
  const DevOpts& o = P.opts;
  const int k = static_cast<int>(P.ix.k);
  int8_t* voteF = reinterpret_cast<int8_t*>(base + P.voteOff);
  int8_t* voteR = voteF + P.pmax;
  const bool useCoverageCheck = o.disableNIP && o.strictCheck;   // include/SACollector.hpp:138
  const bool voteMode = o.strictCheck && !useCoverageCheck;

  for (uint64_t r = static_cast<uint64_t>(blockIdx.x) * WARPS + warp; r < P.reads.numReads; r += static_cast<uint64_t>(gridDim.x) * WARPS) {
    const int mate = r >= P.reads.n ? 1 : 0;
    const uint64_t ri = r - static_cast<uint64_t>(mate) * P.reads.n;
    const uint8_t* src;
    uint32_t len;
    if (P.reads.off[mate]) {
      uint64_t o0 = P.reads.off[mate][ri], o1 = P.reads.off[mate][ri + 1];
      src = P.reads.seq[mate] + o0;
      len = static_cast<uint32_t>(o1 - o0);
    } else {
      src = P.reads.seq[mate] + ri * P.reads.fixedLen;
      len = P.reads.fixedLen;
    }
    ReadSummary s;
    s.ivOff = 0; s.nFwd = 0; s.nRc = 0; s.found = 0; s.pad = 0;
    s.readLen = static_cast<uint16_t>(len);
    if (len > P.maxReadLen) {
      if (lane == 0) { atomicOr(P.status, kStatReadTooLong); s.readLen = 0; P.summ[r] = s; }
      continue;
    }
    const int L = static_cast<int>(len);
    const int npos = L - k + 1;
    __syncwarp();
    bool u = false;
    for (int i = lane; i < L; i += 32) {
      uint8_t ch = __ldg(src + i);
      fwdBuf[i] = upperChar(ch);  // extendSearchNaive compares ::toupper(query); k-mer codes are case-insensitive
      rcBuf[L - 1 - i] = rcChar(ch);
      u |= ((ch | 0x20) == 'u');
    }
    for (int i = lane; i < npos; i += 32) { cF[i] = make_int2(-2, -2); cR[i] = make_int2(-2, -2); }
    const bool hasU = __any_sync(0xffffffffu, u);
    __syncwarp();
    // 2-bit pack both strands + invalid / N masks (32 bases per step: two OR-reductions and two ballots)
    for (uint32_t j = 0; j < pw; ++j) {
      const int i = static_cast<int>(j) * 32 + lane;
#pragma unroll
      for (int strand = 0; strand < 2; ++strand) {
        const uint8_t ch = i < L ? (strand ? rcBuf[i] : fwdBuf[i]) : 0;
        const int cd = baseCode(ch);
        const uint32_t c2 = cd < 0 ? 0u : static_cast<uint32_t>(cd);
        const uint32_t hi = __reduce_or_sync(0xffffffffu, lane < 16 ? (c2 << (30 - 2 * lane)) : 0u);
        const uint32_t lo = __reduce_or_sync(0xffffffffu, lane >= 16 ? (c2 << (30 - 2 * (lane - 16))) : 0u);
        const uint32_t im = __ballot_sync(0xffffffffu, cd < 0);
        const uint32_t nn = __ballot_sync(0xffffffffu, i < L && (ch | 0x20) == 'n');
        if (lane == 0) {
          (strand ? packR : packF)[j] = (static_cast<uint64_t>(hi) << 32) | lo;
          (strand ? invR : invF)[j] = im;
          (strand ? nmR : nmF)[j] = nn;
        }
      }
    }
    __syncwarp();

    WarpCtx& c = *reinterpret_cast<WarpCtx*>(base + P.ctxOff);
    __syncwarp();
    if (lane == 0) {
    c.ix = P.ix; c.fwdBuf = fwdBuf; c.rcBuf = rcBuf; c.cF = cF; c.cR = cR; c.L = L; c.k = k; c.npos = npos; c.hasU = hasU;
    c.packF = packF; c.packR = packR; c.invF = invF; c.invR = invR; c.nmF = nmF; c.nmR = nmR;
    c.maxMMPExtension = o.maxMMPExtension; c.maxInterval = o.maxInterval; c.doChaining = o.doChaining != 0;
    c.disableNIP = o.disableNIP != 0; c.voteMode = voteMode; c.voteF = voteF; c.voteR = voteR;
    }
    if (voteMode) for (int i = lane; i < npos; i += 32) { voteF[i] = 0; voteR[i] = 0; }
    __syncwarp();

    // ---- first-hit scan (SACollector.hpp:167-237)
    uint32_t fwdHit = 0, rcHit = 0, fwdCov = 0, rcCov = 0;
    bool foundHit = false;
    int rb = 0;
    int invalidPos = 0;
    int2 firstIv = make_int2(-1, -1);
    while (rb + k <= L) {
      if (invalidPos != 0x7fffffff) {
        invalidPos = findNMask(nmF, rb, L);
        if (invalidPos <= rb + k) { rb = invalidPos + 1; continue; }  // note <= (SACollector.hpp:178)
      }
      uint64_t mer;
      bool valid = packedKmer(packF, invF, fwdBuf, rb, k, mer);
      if (isHomopolymer(mer, k)) { rb += 1; continue; }
      int2 fm, fc;
      lookupBoth(c, false, rb, mer, valid, true, true, fm, fc);
      if (fm.x >= 0) { ++fwdHit; if (fc.x >= 0) ++rcHit; }
      if (fc.x >= 0 && !fwdHit) ++rcHit;
      if (fwdHit + rcHit > 0) { recordVote(c, false, rb, fm.x >= 0, fc.x >= 0); foundHit = true; firstIv = fm; break; }
      ++rb;
    }
    int nF = 0, nR = 0;
    if (foundHit) {
      bool didCheckFwd = false;
      if (fwdHit) {  // :247-254
        didCheckFwd = true;
        walkStrand(c, false, rb, true, firstIv, fwdCov, fwdHit, rcHit, ivF, nF);
      }
      const bool checkRC = useCoverageCheck ? (rcHit > 0) : (rcHit >= fwdHit);  // :256
      if (checkRC) walkStrand(c, true, 0, false, make_int2(0, 0), rcCov, rcHit, fwdHit, ivR, nR);
      const bool checkFwd = useCoverageCheck ? (fwdHit > 0) : (fwdHit >= rcHit);  // :270
      if (!didCheckFwd && checkFwd) walkStrand(c, false, 0, false, make_int2(0, 0), fwdCov, fwdHit, rcHit, ivF, nF);
      if (useCoverageCheck) {  // strand decision by coverage (:283-288)
        if (fwdCov > rcCov + static_cast<uint32_t>(o.strictCheckSlack)) nR = 0;
        else if (rcCov > fwdCov + static_cast<uint32_t>(o.strictCheckSlack)) nF = 0;
      } else if (o.strictCheck) {  // k-mer "spot check" vote (:289-337)
        if (fwdHit > 0 && rcHit == 0) nR = 0;
        else if (rcHit > 0 && fwdHit == 0) nF = 0;
        else {
          __syncwarp();
          int fs = 0, rs = 0;
          for (int i = lane; i < npos; i += 32) { fs += voteF[i]; rs += voteR[i]; }
          for (int d = 16; d > 0; d >>= 1) { fs += __shfl_xor_sync(0xffffffffu, fs, d); rs += __shfl_xor_sync(0xffffffffu, rs, d); }
          if (fs > rs) nR = 0;
          else if (rs > fs) nF = 0;
        }
      }
      if (o.covReq > 0.0 && o.disableNIP) {  // :343-358
        if (nF > 0 && (static_cast<double>(fwdCov) / static_cast<double>(L)) < o.covReq) nF = 0;
        if (nR > 0 && (static_cast<double>(rcCov) / static_cast<double>(L)) < o.covReq) nR = 0;
      }
    }
    __syncwarp();
    // ---- publish
    const int tot = nF + nR;
    uint32_t off = 0;
    if (tot > 0) {
      if (lane == 0) off = atomicAdd(P.arenaCursor, static_cast<uint32_t>(tot));
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off + static_cast<uint32_t>(tot) > P.arenaCap) {
        if (lane == 0) atomicOr(P.status, kStatIntervalArenaFull);
      } else {
        for (int i = lane; i < tot; i += 32) P.arena[off + i] = i < nF ? ivF[i] : ivR[i - nF];
      }
    }
    if (lane == 0) {
      s.ivOff = off; s.nFwd = static_cast<uint16_t>(nF); s.nRc = static_cast<uint16_t>(nR); s.found = foundHit ? 1 : 0;
      P.summ[r] = s;
    }
  }
}

} // namespace rapmap_b200
