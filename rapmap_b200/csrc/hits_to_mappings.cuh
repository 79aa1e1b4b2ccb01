// Kernel 2 — hit resolution for one read per warp.
//
// Replaces hit_manager::hitsToMappingsSimple (reference src/HitManager.cpp:691-882): the single-interval
// expansion (:716-807), intersectSAHits + intersectSAIntervalWithOutput (:587-689, :449-493),
// collectHitsSimpleSA (:84-326, leftmost anchor or minimap2-style chain DP) and the fwd/rc merge (:834-881).
//
// B200 mapping: the reference walks SA entries one by one into a std::map<tid, ProcessedSAHit>.  Here the
// warp expands all SA entries of a strand at once (lane = entry: one load of the entry's {transcript, position} record,
// 32 in flight), sorts the (tid, interval order, entry) keys with a warp bitonic
// network (shared memory for the common <= 64 entries, an L2-resident global work strip otherwise) and
// resolves each transcript segment with one lane (map iteration order == ascending tid == sorted order).
// Chain scoring keeps the reference's double/float arithmetic with explicit round-to-nearest intrinsics so
// that no FMA contraction can change a tie.
#pragma once
#include "kernels.cuh"

namespace rapmap_b200 {

struct MapParams {
  DeviceIndex ix;
  DevOpts opts;
  uint64_t numReads;
  uint64_t numPairs;       // reads >= numPairs are mate 2
  uint8_t pairedInput;     // 0: unmated reads (MateStatus::SINGLE_END)
  const ReadSummary* summ;
  const IntervalRec* arena;
  QASummary* qsumm;
  QARec* qaArena;
  uint32_t qaCap;
  uint32_t* qaCursor;
  int32_t* posPool;
  uint32_t posCap;
  uint32_t* posCursor;
  uint8_t* scratch;          // per-warp global work strips
  uint32_t scratchEntries;
  uint64_t scratchStride;
  uint32_t smemEntries;
  uint32_t* status;
  uint8_t skipDone;          // 1: reads already resolved by hits_to_mappings_lane_kernel are skipped (nQA != kTodoMark)
  // size-class ordering (k2_class_* kernels): reads grouped by the number of expanded SA entries
  const uint32_t* order;     // nullptr: reads in batch order
  uint32_t* classHist;       // [kK2Buckets] reads per bucket, [kK2Buckets, 2 kK2Buckets) scatter cursors
  uint32_t bLo, bHi;         // buckets this launch works on (inclusive)
};

// bucket = number of expanded SA entries of a read (both strands), 1 .. kK2MaxEntries; 0 = no interval at all;
// kK2MaxEntries + 1 = more entries, or more intervals than a lane kernel takes (warp-per-read kernel)
static constexpr int kK2MaxEntries = 32;
static constexpr int kK2Buckets = kK2MaxEntries + 2;

// [lo, hi) of the positions in P.order that hold the reads of buckets bLo .. bHi (prefix sums of the histogram)
__device__ __forceinline__ void classRange(const MapParams& P, uint32_t& lo, uint32_t& hi) {
  lo = 0; hi = 0;
  for (uint32_t b = 0; b <= P.bHi; ++b) {
    const uint32_t c = P.classHist[b];
    if (b < P.bLo) lo += c;
    hi += c;
  }
}

static constexpr uint32_t kTodoMark = 0xFFFFFFFFu;

// Work-area view (either shared memory or the warp's global strip); every array has `cap` entries.
struct WorkArea {
  uint64_t* keys;    // tid << 32 | ord << 16 | entry ; reused as double f[] by the chain DP
  uint64_t* vals;    // pos << 32 | qpos << 16 | len
  double* qaScore;   // chain score per resolved hit
  QARec* qa;         // resolved hits of this read (fwd list then rc list)
  uint32_t* segs;    // segment starts (bit 31: active)   [cap + 1]
  int32_t* p;        // chain predecessor
  int32_t* aux;      // best chain ends / chain starts
  int32_t* posTmp;   // staged allPositions
  uint8_t* seen;
  uint32_t cap;
};

__host__ __device__ inline uint64_t workAreaBytes(uint32_t entries) {
  uint64_t b = static_cast<uint64_t>(entries) * (8 + 8 + 8 + sizeof(QARec) + 4 + 4 + 4 + 4 + 1) + 32;
  return (b + 15) / 16 * 16;
}

__device__ __forceinline__ WorkArea carve(uint8_t* base, uint32_t entries) {
  WorkArea w;
  w.keys = reinterpret_cast<uint64_t*>(base);
  w.vals = w.keys + entries;
  w.qaScore = reinterpret_cast<double*>(w.vals + entries);
  w.qa = reinterpret_cast<QARec*>(w.qaScore + entries);
  w.segs = reinterpret_cast<uint32_t*>(w.qa + entries);
  w.p = reinterpret_cast<int32_t*>(w.segs + entries + 1);
  w.aux = w.p + entries;
  w.posTmp = w.aux + entries;
  w.seen = reinterpret_cast<uint8_t*>(w.posTmp + entries);
  w.cap = entries;
  return w;
}

__device__ __forceinline__ void warpBitonicSort(uint64_t* keys, uint64_t* vals, int n2, int lane) {
  for (int kk = 2; kk <= n2; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const int lj = __ffs(j) - 1;  // j is a power of two
      for (int t = lane; t < (n2 >> 1); t += 32) {
        int i = ((t >> lj) << (lj + 1)) + (t & (j - 1));
        int l = i + j;
        bool up = ((i & kk) == 0);
        uint64_t a = keys[i], b = keys[l];
        if ((a > b) == up) {
          keys[i] = b; keys[l] = a;
          uint64_t va = vals[i]; vals[i] = vals[l]; vals[l] = va;
        }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ uint32_t vPos(uint64_t v) { return static_cast<uint32_t>(v >> 32); }
__device__ __forceinline__ uint32_t vQpos(uint64_t v) { return static_cast<uint32_t>(v >> 16) & 0xffffu; }
__device__ __forceinline__ int32_t vLen(uint64_t v) { return static_cast<int32_t>(v & 0xffffu); }

// fastlog2 (reference src/HitManager.cpp:33-42), float ops in source order, never contracted.
__device__ __forceinline__ float fastlog2Ref(float x) {
  uint32_t vi = __float_as_uint(x);
  float mx = __uint_as_float((vi & 0x007FFFFFu) | 0x3f000000u);
  float y = __uint2float_rn(vi);
  y = __fmul_rn(y, 1.1920928955078125e-7f);
  float t = __fsub_rn(y, 124.22551499f);
  t = __fsub_rn(t, __fmul_rn(1.498030302f, mx));
  t = __fsub_rn(t, __fdiv_rn(1.72587999f, __fadd_rn(0.3520887068f, mx)));
  return t;
}

// alpha / beta of collectHitsSimpleSA (:126-139); returns f[j] + alpha - beta.
__device__ __forceinline__ double extensionScore(double fj, int32_t qdiff, int32_t rdiff, int32_t ilen, int32_t maxDist) {
  double score = static_cast<double>(ilen);
  double mindiff = static_cast<double>((qdiff < rdiff) ? qdiff : rdiff);
  double alpha = (score < mindiff) ? score : mindiff;
  double beta;
  if (qdiff < 0 || ((qdiff > rdiff ? qdiff : rdiff) > maxDist)) {
    beta = __longlong_as_double(0x7ff0000000000000LL);
  } else {
    double l = static_cast<double>(qdiff - rdiff);
    int32_t al = static_cast<int32_t>(fabs(l));
    if (l == 0.0) beta = 0.0;
    else beta = __dadd_rn(__dmul_rn(__dmul_rn(0.01, 31.0), static_cast<double>(al)), __dmul_rn(0.5, static_cast<double>(fastlog2Ref(static_cast<float>(al)))));
  }
  return __dsub_rn(__dadd_rn(fj, alpha), beta);
}

// One transcript segment [b, e) of sorted entries, chain mode (collectHitsSimpleSA :107-306).  Executed by one lane.
// Writes the hit positions (allPositions) to w.posTmp[b ...] and returns their count; q gets pos / chain status.
__device__ __forceinline__ uint32_t chainSegment(WorkArea& w, uint32_t b, uint32_t e, uint32_t readLen, int32_t maxDist, bool considerMultiPos,
                                                 QARec& q, double& bestScoreOut) {
  const int32_t m = static_cast<int32_t>(e - b);
  uint64_t* v = w.vals + b;
  double* f = reinterpret_cast<double*>(w.keys + b);
  int32_t* p = w.p + b;
  int32_t* ends = w.aux + b;
  uint8_t* seen = w.seen + b;
  // sort by (ref end, query end) — keys are unique within a transcript (:111-121), insertion sort of a short list
  for (int32_t i = 1; i < m; ++i) {
    uint64_t x = v[i];
    uint32_t xr = vPos(x) + static_cast<uint32_t>(vLen(x)), xq = vQpos(x) + static_cast<uint32_t>(vLen(x));
    int32_t j = i - 1;
    while (j >= 0) {
      uint64_t y = v[j];
      uint32_t yr = vPos(y) + static_cast<uint32_t>(vLen(y)), yq = vQpos(y) + static_cast<uint32_t>(vLen(y));
      bool less = (xr < yr) || (xr == yr && xq < yq);
      if (!less) break;
      v[j + 1] = y;
      --j;
    }
    v[j + 1] = x;
  }
  double bestScore = -1.7976931348623157e308;  // numeric_limits<double>::lowest()
  int32_t bestChainEnd = -1;
  int32_t nEnds = 0;
  for (int32_t i = 0; i < m; ++i) {
    uint64_t hi = v[i];
    const int32_t len = vLen(hi);
    const uint32_t qposi = vQpos(hi) + static_cast<uint32_t>(len), rposi = vPos(hi) + static_cast<uint32_t>(len);
    int32_t pi = i;
    double fi = static_cast<double>(len);
    int32_t numRounds = 2;
    for (int32_t j = i - 1; j >= 0; --j) {
      uint64_t hj = v[j];
      const uint32_t qposj = vQpos(hj) + static_cast<uint32_t>(vLen(hj)), rposj = vPos(hj) + static_cast<uint32_t>(vLen(hj));
      const int32_t qdiff = static_cast<int32_t>(qposi - qposj), rdiff = static_cast<int32_t>(rposi - rposj);
      double ext = extensionScore(f[j], qdiff, rdiff, len, maxDist);
      bool extendWithJ = ext > fi;
      pi = extendWithJ ? j : pi;
      fi = extendWithJ ? ext : fi;
      if (pi < i) { numRounds--; if (numRounds <= 0) break; }
    }
    p[i] = pi;
    f[i] = fi;
    seen[i] = 0;
    if (fi > bestScore) {
      bestScore = fi; bestChainEnd = i;
      if (considerMultiPos) { nEnds = 0; ends[nEnds++] = i; }
    } else if (considerMultiPos && fi == bestScore) {
      ends[nEnds++] = i;
    }
  }
  if (!considerMultiPos) { nEnds = 0; ends[nEnds++] = bestChainEnd; }
  // backtracking with the seen[] de-duplication (:220-257); chain starts overwrite ends[] in place
  int32_t nStarts = 0;
  uint32_t numDistinctOpt = 0;
  for (int32_t k = 0; k < nEnds; ++k) {
    int32_t cur = ends[k];
    bool valid = true;
    int32_t lastPtr = p[cur];
    while (lastPtr < cur) {
      if (seen[cur]) { valid = false; break; }
      seen[cur] = 1;
      cur = lastPtr;
      lastPtr = p[cur];
    }
    if (seen[cur]) valid = false;
    if (valid) { ++numDistinctOpt; ends[nStarts++] = lastPtr; }
  }
  // hit record (:259-280)
  int32_t* out = w.posTmp + b;
  for (int32_t k = 0; k < nStarts; ++k) {
    uint64_t s = v[ends[k]];
    out[k] = static_cast<int32_t>(vPos(s) - vQpos(s));
  }
  q.pos = out[0];
  if (nStarts > 1) {  // std::sort(allPositions)
    for (int32_t i = 1; i < nStarts; ++i) {
      int32_t x = out[i];
      int32_t j = i - 1;
      while (j >= 0 && out[j] > x) { out[j + 1] = out[j]; --j; }
      out[j + 1] = x;
    }
  }
  // gapless chain spanning the read => UNGAPPED (:283-306)
  q.chain = 4;
  if (m > 1 && numDistinctOpt == 1 && bestChainEnd == m - 1) {
    uint64_t last = v[m - 1], first = v[0];
    int64_t queryRange = static_cast<int64_t>(vQpos(last) + static_cast<uint32_t>(vLen(last))) - static_cast<int64_t>(vQpos(first));
    int64_t refRange = static_cast<int64_t>(vPos(last) + static_cast<uint32_t>(vLen(last))) - static_cast<int64_t>(vPos(first));
    if (queryRange == refRange && queryRange == static_cast<int64_t>(readLen)) q.chain = 1;
  }
  bestScoreOut = bestScore;
  return static_cast<uint32_t>(nStarts);
}

// Resolves one strand: appends QARecs (ascending tid) to w.qa[nOut...], their positions to w.posTmp.
#ifndef RAPMAP_K2_STRAND_ATTR
#define RAPMAP_K2_STRAND_ATTR __noinline__  // one copy of the strand code: the inlined pair thrashed the instruction cache (no_instruction stalls, profiles/r01g)
#endif
__device__ RAPMAP_K2_STRAND_ATTR void resolveStrand(const MapParams& P, WorkArea& w, const IntervalRec* ivs, int nIv, bool isFw, uint32_t readLen,
                                              int lane, uint32_t& nOut, uint32_t& posBase) {
  const DeviceIndex& ix = P.ix;
  const DevOpts& o = P.opts;
  const bool needPos = o.selAln || o.fuzzy;
  // ---- processing order: smallest-span interval first (first wins on ties, :636-641), others in original order
  int minIdx = 0;
  uint32_t total = 0;
  {
    int bestSpan = 0x7fffffff;
    for (int j = 0; j < nIv; ++j) {
      int span = ivs[j].end - ivs[j].begin;
      total += static_cast<uint32_t>(span);
      if (nIv > 1 && span < bestSpan) { bestSpan = span; minIdx = j; }
    }
  }
  // ---- expand: lane = SA entry.  Entries of this strand occupy [posBase, posBase + total) of the work arrays.
  uint64_t* keys = w.keys + posBase;
  uint64_t* vals = w.vals + posBase;
  uint32_t base = 0;
  for (int j = 0; j < nIv; ++j) {
    IntervalRec iv = ivs[j];
    uint32_t ord = (nIv == 1) ? 0u : (j == minIdx ? 0u : static_cast<uint32_t>(j < minIdx ? j + 1 : j));
    int span = iv.end - iv.begin;
    for (int e = lane; e < span; e += 32) {
      const uint2 tp = __ldg(ix.saTidPos + iv.begin + e);  // {transcript, position in it} of SA[begin + e]
      const uint32_t tid = tp.x;
      const int32_t pos = static_cast<int32_t>(tp.y);
      keys[base + e] = (static_cast<uint64_t>(tid) << 32) | (static_cast<uint64_t>(ord) << 16) | static_cast<uint64_t>(e);
      vals[base + e] = (static_cast<uint64_t>(static_cast<uint32_t>(pos)) << 32) | (static_cast<uint64_t>(iv.qpos) << 16) | iv.len;
    }
    base += static_cast<uint32_t>(span);
  }
  __syncwarp();
  // ---- sort by (tid, ord, entry)
  if (total > 1 && total <= 32) {
    // rank by counting: every lane holds one key, its rank is the number of smaller keys (keys are unique)
    const bool mine = static_cast<uint32_t>(lane) < total;
    const uint64_t kx = mine ? keys[lane] : ~0ULL;
    const uint64_t vx = mine ? vals[lane] : 0ULL;
    uint32_t rank = 0;
    for (uint32_t j = 0; j < total; ++j) {
      const uint64_t kj = __shfl_sync(0xffffffffu, kx, j);
      rank += kj < kx ? 1u : 0u;
    }
    __syncwarp();
    if (mine) { keys[rank] = kx; vals[rank] = vx; }
    __syncwarp();
  } else if (total > 1) {
    int n2 = 1;
    while (n2 < static_cast<int>(total)) n2 <<= 1;
    for (int i = total + lane; i < n2; i += 32) { keys[i] = ~0ULL; vals[i] = 0; }
    __syncwarp();
    warpBitonicSort(keys, vals, n2, lane);
  }
  // ---- segment heads
  uint32_t* segs = w.segs;
  uint32_t nSeg = 0;
  for (uint32_t b0 = 0; b0 < total; b0 += 32) {
    uint32_t i = b0 + lane;
    bool head = false;
    if (i < total) head = (i == 0) || ((keys[i] >> 32) != (keys[i - 1] >> 32));
    unsigned mk = __ballot_sync(0xffffffffu, head);
    if (head) segs[nSeg + __popc(mk & ((1u << lane) - 1u))] = i;
    nSeg += __popc(mk);
  }
  if (lane == 0) segs[nSeg] = total;
  __syncwarp();
  // ---- which transcripts are active (:613-628, :667-686)
  int32_t required = nIv, maxSlack = 0;
  if (nIv > 1 && o.consensusFraction < 1.0f) {
    float requiredFrac = __fmul_rn(static_cast<float>(nIv), o.consensusFraction);
    int32_t fl = static_cast<int32_t>(floorf(requiredFrac));
    required = fl > 1 ? fl : 1;
    maxSlack = nIv - required;
  }
  if (nIv > 1) {
    bool anyActive = false;
    for (uint32_t s0 = 0; s0 < nSeg; s0 += 32) {
      uint32_t s = s0 + lane;
      bool act = false;
      if (s < nSeg) {
        uint32_t b = segs[s], e = segs[s + 1] & 0x7fffffffu;
        int32_t distinct = 0;
        uint32_t lastOrd = 0xffffffffu;
        for (uint32_t i = b; i < e; ++i) {
          uint32_t ord = static_cast<uint32_t>(keys[i] >> 16) & 0xffffu;
          if (ord != lastOrd) { ++distinct; lastOrd = ord; }
        }
        // strict intersection (maxSlack == 0): a transcript survives only if present in every interval; with slack
        // every entry is recorded and numActive is the number of distinct intervals (:460,:473-490)
        act = distinct >= required;
      }
      anyActive |= __any_sync(0xffffffffu, act);
      __syncwarp();
      if (s < nSeg && act) segs[s] |= 0x80000000u;
      __syncwarp();
    }
    if (maxSlack > 0 && !anyActive) {
      for (uint32_t s = lane; s < nSeg; s += 32) segs[s] |= 0x80000000u;
      __syncwarp();
    }
  }
  // ---- one lane per transcript segment
  const bool chain = o.doChaining && nIv > 1;
  for (uint32_t s0 = 0; s0 < nSeg; s0 += 32) {
    uint32_t s = s0 + lane;
    bool active = false;
    QARec q;
    double score = -1.7976931348623157e308;
    if (s < nSeg) {
      uint32_t sb = segs[s];
      uint32_t b = sb & 0x7fffffffu, e = segs[s + 1] & 0x7fffffffu;
      active = (nIv == 1) || (sb & 0x80000000u);
      if (active) {
        q.tid = static_cast<uint32_t>(keys[b] >> 32);
        q.fwd = isFw ? 1 : 0; q.pad = 0; q.oppOff = 0; q.nOpp = 0; q.posOff = posBase + b; q.nAll = 1; q.chain = 4;
        if (chain) {
          q.nAll = chainSegment(w, posBase + b, posBase + e, readLen, static_cast<int32_t>(readLen), o.considerMultiPos, q, score);
        } else if (nIv == 1) {
          // collectFromSingleInterval (:716-807): sorted by (tid, pos); first position is the hit, the rest are allPositions
          q.chain = (ivs[0].len == readLen) ? 0 : 4;  // PERFECT iff the single MMP spans the read (:741-746)
          const int32_t qp = static_cast<int32_t>(ivs[0].qpos);
          if (o.considerMultiPos) {
            int32_t* out = w.posTmp + posBase + b;
            uint32_t cnt = 0;
            for (uint32_t i = b; i < e; ++i) {
              int32_t hp = static_cast<int32_t>(vPos(vals[i])) - qp;
              int32_t j = static_cast<int32_t>(cnt) - 1;
              while (j >= 0 && out[j] > hp) { out[j + 1] = out[j]; --j; }
              out[j + 1] = hp;
              ++cnt;
            }
            q.pos = out[0]; q.nAll = cnt;
          } else {
            int32_t best = 0x7fffffff;
            for (uint32_t i = b; i < e; ++i) { int32_t hp = static_cast<int32_t>(vPos(vals[i])) - qp; best = hp < best ? hp : best; }
            q.pos = best;
            if (needPos) w.posTmp[posBase + b] = best;
          }
        } else {
          // leftmost anchor (:308-322): std::min_element keeps the first minimum in tqvec order == (ord, entry) order
          uint32_t bestPos = 0xffffffffu;
          int32_t bestHit = 0;
          for (uint32_t i = b; i < e; ++i) {
            uint64_t v = vals[i];
            if (vPos(v) < bestPos) { bestPos = vPos(v); bestHit = static_cast<int32_t>(vPos(v) - vQpos(v)); }
          }
          q.pos = bestHit;
          if (needPos) w.posTmp[posBase + b] = bestHit;
        }
      }
    }
    unsigned mk = __ballot_sync(0xffffffffu, active);
    if (active) {
      uint32_t slot = nOut + __popc(mk & ((1u << lane) - 1u));
      w.qa[slot] = q;
      w.qaScore[slot] = score;
    }
    nOut += __popc(mk);
  }
  posBase += total;
  __syncwarp();
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) hits_to_mappings_kernel(MapParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint64_t gw = static_cast<uint64_t>(blockIdx.x) * WARPS + warp;
  uint8_t* smemBase = smem + static_cast<size_t>(warp) * workAreaBytes(P.smemEntries);
  uint8_t* globBase = P.scratch + gw * P.scratchStride;
  const bool needPos = P.opts.selAln || P.opts.fuzzy;

  uint32_t pLo = 0, pHi = static_cast<uint32_t>(P.numReads);
  if (P.order) classRange(P, pLo, pHi);
  for (uint64_t p = pLo + gw; p < pHi; p += static_cast<uint64_t>(gridDim.x) * WARPS) {
    const uint64_t r = P.order ? P.order[p] : p;
    if (P.skipDone && P.qsumm[r].nQA != kTodoMark) continue;
    ReadSummary s = P.summ[r];
    QASummary out;
    out.qaOff = 0; out.nQA = 0;
    const int nF = s.nFwd, nR = s.nRc;
    if (nF + nR == 0) {
      if (lane == 0) P.qsumm[r] = out;
      continue;
    }
    const IntervalRec* ivs = P.arena + s.ivOff;
    // work-area demand: the two strands sit side by side; each is padded to a power of two for the sort, and the
    // fwd/rc merge writes its result behind the two hit lists (<= 2 * entries)
    uint32_t totF = 0, totR = 0;
    for (int j = 0; j < nF; ++j) totF += static_cast<uint32_t>(ivs[j].end - ivs[j].begin);
    for (int j = nF; j < nF + nR; ++j) totR += static_cast<uint32_t>(ivs[j].end - ivs[j].begin);
    uint32_t padR = 1;
    while (padR < totR) padR <<= 1;
    uint32_t padF = 1;
    while (padF < totF) padF <<= 1;
    uint32_t need = totF + (padR > padF ? padR : padF);
    if (nF > 0 && nR > 0) need = need > 2 * (totF + totR) ? need : 2 * (totF + totR);
    WorkArea w;
    if (need <= P.smemEntries) w = carve(smemBase, P.smemEntries);
    else if (need <= P.scratchEntries) w = carve(globBase, P.scratchEntries);
    else {
      if (lane == 0) { atomicOr(P.status, kStatScratchFull); P.qsumm[r] = out; }
      continue;
    }
    uint32_t nOut = 0, posBase = 0;
    if (nF > 0) resolveStrand(P, w, ivs, nF, true, s.readLen, lane, nOut, posBase);
    const uint32_t nFwdOut = nOut;
    if (nR > 0) resolveStrand(P, w, ivs + nF, nR, false, s.readLen, lane, nOut, posBase);
    uint32_t nFinal = nOut;
    // ---- merge forward and reverse-complement lists (:834-881).  Rare (both strands kept), done by lane 0:
    // stable merge on tid, equal tid => higher chain score first (forward first on a tie); the survivor takes the
    // loser's allPositions as oppositeStrandPositions.
    if (nFwdOut > 0 && nOut > nFwdOut) {
      if (lane == 0) {
        uint32_t a = 0, b = nFwdOut, oo = nOut;
        while (a < nFwdOut || b < nOut) {
          bool takeA;
          if (b >= nOut) takeA = true;
          else if (a >= nFwdOut) takeA = false;
          else if (w.qa[a].tid != w.qa[b].tid) takeA = w.qa[a].tid < w.qa[b].tid;
          else {
            // same transcript on both strands
            bool rcFirst = w.qaScore[b] > w.qaScore[a];
            QARec win = rcFirst ? w.qa[b] : w.qa[a];
            const QARec& lose = rcFirst ? w.qa[a] : w.qa[b];
            win.oppOff = lose.posOff; win.nOpp = lose.nAll;
            w.qaScore[oo] = rcFirst ? w.qaScore[b] : w.qaScore[a];
            w.qa[oo++] = win;
            ++a; ++b;
            continue;
          }
          if (takeA) { w.qaScore[oo] = w.qaScore[a]; w.qa[oo++] = w.qa[a++]; }
          else { w.qaScore[oo] = w.qaScore[b]; w.qa[oo++] = w.qa[b++]; }
        }
        nFinal = oo - nOut;
        for (uint32_t i = 0; i < nFinal; ++i) w.qa[i] = w.qa[nOut + i];
      }
      nFinal = __shfl_sync(0xffffffffu, nFinal, 0);
      __syncwarp();
    }
    // ---- publish hits (+ position lists)
    uint32_t off = 0;
    bool ok = true;
    if (nFinal > 0) {
      if (lane == 0) off = atomicAdd(P.qaCursor, nFinal);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off + nFinal > P.qaCap) {
        if (lane == 0) atomicOr(P.status, kStatQAArenaFull);
        ok = false;
      }
      uint32_t poolOff = 0, nPosTot = 0;
      if (needPos) {
        for (uint32_t i = lane; i < nFinal; i += 32) nPosTot += w.qa[i].nAll + w.qa[i].nOpp;
        for (int d = 16; d > 0; d >>= 1) nPosTot += __shfl_xor_sync(0xffffffffu, nPosTot, d);
        if (lane == 0) poolOff = atomicAdd(P.posCursor, nPosTot);
        poolOff = __shfl_sync(0xffffffffu, poolOff, 0);
        if (poolOff + nPosTot > P.posCap) {
          if (lane == 0) atomicOr(P.status, kStatPosPoolFull);
          ok = false;
        }
      }
      if (ok && !needPos) {
        for (uint32_t i = lane; i < nFinal; i += 32) P.qaArena[off + i] = w.qa[i];
      } else if (ok) {
        uint32_t run = poolOff;
        for (uint32_t i = 0; i < nFinal; ++i) {  // warp walks the hits, lanes copy the position lists
          QARec q = w.qa[i];
          if (needPos) {
            for (uint32_t j = lane; j < q.nAll; j += 32) P.posPool[run + j] = w.posTmp[q.posOff + j];
            for (uint32_t j = lane; j < q.nOpp; j += 32) P.posPool[run + q.nAll + j] = w.posTmp[q.oppOff + j];
            q.posOff = run; q.oppOff = run + q.nAll;
            run += q.nAll + q.nOpp;
          }
          if (lane == 0) P.qaArena[off + i] = q;
        }
      } else {
        nFinal = 0;
      }
    }
    if (lane == 0) { out.qaOff = off; out.nQA = nFinal; P.qsumm[r] = out; }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Lane-per-read form for the common small case: no chaining / position lists (plain quasimap), at most LANE_MAXIV
// intervals and CAP expanded SA entries per read.  One thread resolves one read: the {transcript, position} records of ALL
// its SA entries are loaded together, so a lane has up to CAP loads in flight and a warp 32 reads; the (tid, interval order, entry) keys are insertion-sorted in the lane's
// shared-memory strip (interleaved across the block), segments are resolved with the same rules as resolveStrand
// (strict / consensus intersection, leftmost anchor, PERFECT flag of a single read-spanning interval), the two strands
// are merged by tid.  Anything bigger is marked kTodoMark and left to the warp-per-read kernel.
static constexpr int kLaneMaxIv = 8;

template <int NT, int CAP>
__global__ void __launch_bounds__(NT) hits_to_mappings_lane_kernel(MapParams P) {
  extern __shared__ __align__(16) uint8_t smemLaneRaw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smemLaneRaw) + threadIdx.x;  // entry i at keys[i * NT]
  uint64_t* vals = keys + CAP * NT;
  const int lane = threadIdx.x & 31;
  const DeviceIndex& ix = P.ix;
  const DevOpts& o = P.opts;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * NT;
  for (uint64_t base = static_cast<uint64_t>(blockIdx.x) * NT + (threadIdx.x & ~31); base < P.numReads; base += stride) {
    const uint64_t r = base + lane;
    const bool valid = r < P.numReads;
    int nF = 0, nR = 0;
    uint32_t ivOff = 0, readLen = 0;
    if (valid) {
      const ReadSummary s = P.summ[r];
      nF = s.nFwd; nR = s.nRc; ivOff = s.ivOff; readLen = s.readLen;
    }
    const int nIv = nF + nR;
    bool todo = false;
    uint32_t totF = 0, totR = 0;
    const IntervalRec* ivs = P.arena + ivOff;
    if (nIv > kLaneMaxIv) todo = true;
    else {
      for (int j = 0; j < nIv; ++j) {
        const uint32_t span = static_cast<uint32_t>(ivs[j].end - ivs[j].begin);
        if (j < nF) totF += span; else totR += span;
      }
      if (totF + totR > static_cast<uint32_t>(CAP)) todo = true;
    }
    const uint32_t total = todo ? 0u : totF + totR;
    // ---- stage 1: one key/value per SA entry (forward strand first); value carries the SA index for now
    if (total > 0) {
      uint32_t at = 0;
      for (int strand = 0; strand < 2; ++strand) {
        const int j0 = strand ? nF : 0, j1 = strand ? nIv : nF, n = j1 - j0;
        int minIdx = 0, bestSpan = 0x7fffffff;  // smallest-span interval is processed first (first wins on ties, :636-641)
        for (int j = 0; j < n; ++j) {
          const int span = ivs[j0 + j].end - ivs[j0 + j].begin;
          if (n > 1 && span < bestSpan) { bestSpan = span; minIdx = j; }
        }
        for (int j = 0; j < n; ++j) {
          const IntervalRec iv = ivs[j0 + j];
          const uint32_t ord = (n == 1) ? 0u : (j == minIdx ? 0u : static_cast<uint32_t>(j < minIdx ? j + 1 : j));
          const int span = iv.end - iv.begin;
          for (int e = 0; e < span; ++e) {
            keys[at * NT] = (static_cast<uint64_t>(ord) << 16) | static_cast<uint64_t>(e);
            vals[at * NT] = (static_cast<uint64_t>(static_cast<uint32_t>(iv.begin + e)) << 32) | (static_cast<uint64_t>(iv.qpos) << 16) | iv.len;
            ++at;
          }
        }
      }
    }
    // ---- stage 2: SA entry -> {transcript, position in the transcript}: one independent load per entry (the image holds the
    //      pair per SA entry; it used to be the dependent chain SA -> rank record -> txpOffsets)
#pragma unroll 4
    for (uint32_t i = 0; i < total; ++i) {
      const uint64_t v = vals[i * NT];
      const uint2 tp = __ldg(ix.saTidPos + (v >> 32));
      keys[i * NT] |= static_cast<uint64_t>(tp.x) << 32;
      vals[i * NT] = (static_cast<uint64_t>(tp.y) << 32) | (v & 0xffffffffULL);
    }
    // ---- per strand: sort by (tid, ord, entry), resolve the transcript segments, compact {tid, pos, chain} in place
    uint32_t nOutF = 0, nOutR = 0;
    for (int strand = 0; strand < 2 && total > 0; ++strand) {
      const uint32_t b0 = strand ? totF : 0u, cnt = strand ? totR : totF;
      const int n = strand ? nR : nF;
      if (cnt == 0) continue;
      for (uint32_t i = 1; i < cnt; ++i) {  // insertion sort (keys are unique)
        const uint64_t kx = keys[(b0 + i) * NT], vx = vals[(b0 + i) * NT];
        int32_t j = static_cast<int32_t>(i) - 1;
        while (j >= 0 && keys[(b0 + j) * NT] > kx) {
          keys[(b0 + j + 1) * NT] = keys[(b0 + j) * NT];
          vals[(b0 + j + 1) * NT] = vals[(b0 + j) * NT];
          --j;
        }
        keys[(b0 + j + 1) * NT] = kx; vals[(b0 + j + 1) * NT] = vx;
      }
      int32_t required = n, maxSlack = 0;  // :613-628
      if (n > 1 && o.consensusFraction < 1.0f) {
        const float requiredFrac = __fmul_rn(static_cast<float>(n), o.consensusFraction);
        const int32_t fl = static_cast<int32_t>(floorf(requiredFrac));
        required = fl > 1 ? fl : 1;
        maxSlack = n - required;
      }
      const uint32_t ivLen0 = ivs[strand ? nF : 0].len, ivQpos0 = ivs[strand ? nF : 0].qpos;
      uint32_t nOut = 0;
      for (int pass = 0; pass < 2; ++pass) {  // pass 1 only when the consensus rule leaves no transcript: then all are kept (:667-686)
        const bool keepAll = pass == 1;
        uint32_t i = 0;
        nOut = 0;
        while (i < cnt) {
          const uint32_t tid = static_cast<uint32_t>(keys[(b0 + i) * NT] >> 32);
          uint32_t e = i;
          int32_t distinct = 0;
          uint32_t lastOrd = 0xffffffffu, bestPos = 0xffffffffu;
          int32_t leftmost = 0, minHit = 0x7fffffff;
          while (e < cnt && static_cast<uint32_t>(keys[(b0 + e) * NT] >> 32) == tid) {
            const uint32_t ord = static_cast<uint32_t>(keys[(b0 + e) * NT] >> 16) & 0xffffu;
            if (ord != lastOrd) { ++distinct; lastOrd = ord; }
            const uint64_t v = vals[(b0 + e) * NT];
            if (vPos(v) < bestPos) { bestPos = vPos(v); leftmost = static_cast<int32_t>(vPos(v) - vQpos(v)); }  // leftmost anchor (:308-322)
            const int32_t hp = static_cast<int32_t>(vPos(v)) - static_cast<int32_t>(ivQpos0);
            minHit = hp < minHit ? hp : minHit;
            ++e;
          }
          const bool active = (n == 1) || keepAll || distinct >= required;
          if (active && pass == (keepAll ? 1 : 0)) {
            // single interval (:716-807): smallest position, PERFECT iff the MMP spans the read (:741-746)
            const uint32_t chain = (n == 1 && ivLen0 == readLen) ? 0u : 4u;
            const int32_t pos = (n == 1) ? minHit : leftmost;
            // in-place: nOut <= i, so the write never overtakes the scan; the scan of this segment is already done
            keys[(b0 + nOut) * NT] = (static_cast<uint64_t>(tid) << 32) | chain;
            vals[(b0 + nOut) * NT] = static_cast<uint64_t>(static_cast<uint32_t>(pos));
            ++nOut;
          }
          i = e;
        }
        if (nOut > 0 || !(n > 1 && maxSlack > 0)) break;
        // consensus mode and nothing active: every transcript is kept.  The in-place compaction wrote nothing (nOut == 0),
        // so the sorted entries are intact for the second pass.
      }
      if (strand) nOutR = nOut; else nOutF = nOut;
    }
    // ---- merge by tid (:834-881; without chain scores the forward hit wins a shared transcript) and publish
    uint32_t nFinal = 0;
    if (total > 0) {
      uint32_t a = 0, b = 0;
      while (a < nOutF || b < nOutR) {
        if (b >= nOutR) { ++a; }
        else if (a >= nOutF) { ++b; }
        else {
          const uint32_t ta = static_cast<uint32_t>(keys[a * NT] >> 32), tb = static_cast<uint32_t>(keys[(totF + b) * NT] >> 32);
          if (ta == tb) { ++a; ++b; } else if (ta < tb) ++a; else ++b;
        }
        ++nFinal;
      }
    }
    const unsigned pm = __ballot_sync(0xffffffffu, nFinal > 0);
    uint32_t off = 0;
    if (pm) {
      int incl = static_cast<int>(nFinal);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const uint32_t tot = static_cast<uint32_t>(__shfl_sync(0xffffffffu, incl, 31));
      uint32_t wbase = 0;
      if (lane == 0) wbase = atomicAdd(P.qaCursor, tot);
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      off = wbase + static_cast<uint32_t>(incl) - nFinal;
      if (static_cast<uint64_t>(wbase) + tot > P.qaCap) {
        if (lane == 0) atomicOr(P.status, kStatQAArenaFull);
        nFinal = 0;
      }
    }
    if (nFinal > 0) {
      uint32_t a = 0, b = 0, w = 0;
      while (a < nOutF || b < nOutR) {
        bool takeA;
        if (b >= nOutR) takeA = true;
        else if (a >= nOutF) takeA = false;
        else {
          const uint32_t ta = static_cast<uint32_t>(keys[a * NT] >> 32), tb = static_cast<uint32_t>(keys[(totF + b) * NT] >> 32);
          if (ta == tb) { takeA = true; ++b; } else takeA = ta < tb;
        }
        const uint32_t src = takeA ? a : totF + b;
        const uint64_t kx = keys[src * NT];
        QARec q;
        q.tid = static_cast<uint32_t>(kx >> 32);
        q.pos = static_cast<int32_t>(static_cast<uint32_t>(vals[src * NT]));
        q.posOff = 0; q.nAll = 1; q.oppOff = 0; q.nOpp = 0;
        q.fwd = takeA ? 1 : 0; q.chain = static_cast<uint8_t>(kx & 0xffu); q.pad = 0;
        P.qaArena[off + w] = q;
        ++w;
        if (takeA) ++a; else ++b;
      }
    }
    if (valid) {
      QASummary out;
      out.qaOff = off; out.nQA = todo ? kTodoMark : nFinal;
      P.qsumm[r] = out;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Lane-per-read form WITH chaining / position lists (quasimap -s, -f): the warp-per-read kernel gives one lane to each
// transcript segment of ONE read, so the chain DP of a read with 4 transcripts runs on 4 of 32 threads (17 active
// threads per instruction, 34 % issue, profiles/r01h).  Here a thread owns a read with at most kChainLaneMaxIv intervals
// and CAP expanded SA entries; its work arrays (keys / values / DP links / positions) are a private strip of shared
// memory, so chainSegment() is used unchanged.  Loads are staged for all entries of the read as in the plain lane
// kernel.  Everything bigger is marked kTodoMark for the warp-per-read kernel.
static constexpr int kChainLaneMaxIv = 16;

__host__ __device__ inline uint32_t chainLaneStride(uint32_t cap) { return (cap * 33u + 8u + 7u) / 8u * 8u; }

template <int NT, int CAP>
__global__ void __launch_bounds__(NT) hits_to_mappings_chain_lane_kernel(MapParams P) {
  extern __shared__ __align__(16) uint8_t smemChain[];
  uint8_t* base = smemChain + static_cast<size_t>(threadIdx.x) * chainLaneStride(CAP);
  WorkArea w;
  w.keys = reinterpret_cast<uint64_t*>(base);
  w.vals = w.keys + CAP;
  w.posTmp = reinterpret_cast<int32_t*>(w.vals + CAP);
  w.p = w.posTmp + CAP;
  w.aux = w.p + CAP;
  uint32_t* outMeta = reinterpret_cast<uint32_t*>(w.aux + CAP);  // per resolved hit: posOff << 16 | nAll << 8 | chain
  w.seen = reinterpret_cast<uint8_t*>(outMeta + CAP);
  w.qaScore = nullptr; w.qa = nullptr; w.segs = nullptr; w.cap = CAP;
  uint64_t* keys = w.keys;
  uint64_t* vals = w.vals;
  const int lane = threadIdx.x & 31;
  const DeviceIndex& ix = P.ix;
  const DevOpts& o = P.opts;
  const bool needPos = o.selAln || o.fuzzy;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * NT;
  uint32_t pLo = 0, pHi = static_cast<uint32_t>(P.numReads);
  if (P.order) classRange(P, pLo, pHi);  // the reads of this launch's size range, contiguous in P.order
  for (uint64_t rbase = pLo + static_cast<uint64_t>(blockIdx.x) * NT + (threadIdx.x & ~31); rbase < pHi; rbase += stride) {
    const bool valid = rbase + lane < pHi;
    const uint64_t r = valid ? (P.order ? P.order[rbase + lane] : rbase + lane) : 0;
    int nF = 0, nR = 0;
    uint32_t ivOff = 0, readLen = 0;
    if (valid) {
      const ReadSummary s = P.summ[r];
      nF = s.nFwd; nR = s.nRc; ivOff = s.ivOff; readLen = s.readLen;
    }
    const int nIv = nF + nR;
    bool todo = false;
    uint32_t totF = 0, totR = 0;
    const IntervalRec* ivs = P.arena + ivOff;
    if (nIv > kChainLaneMaxIv) todo = true;
    else {
      for (int j = 0; j < nIv; ++j) {
        const uint32_t span = static_cast<uint32_t>(ivs[j].end - ivs[j].begin);
        if (j < nF) totF += span; else totR += span;
      }
      if (totF + totR > static_cast<uint32_t>(CAP)) todo = true;
    }
    if (todo && P.order) atomicOr(P.status, kStatInternal);  // the size classes promise reads that fit the strip
    const uint32_t total = todo ? 0u : totF + totR;
    // ---- stage 1: one key/value per SA entry (forward strand first); the value carries the SA index for now
    if (total > 0) {
      uint32_t at = 0;
      for (int strand = 0; strand < 2; ++strand) {
        const int j0 = strand ? nF : 0, j1 = strand ? nIv : nF, n = j1 - j0;
        int minIdx = 0, bestSpan = 0x7fffffff;
        for (int j = 0; j < n; ++j) {
          const int span = ivs[j0 + j].end - ivs[j0 + j].begin;
          if (n > 1 && span < bestSpan) { bestSpan = span; minIdx = j; }
        }
        for (int j = 0; j < n; ++j) {
          const IntervalRec iv = ivs[j0 + j];
          const uint32_t ord = (n == 1) ? 0u : (j == minIdx ? 0u : static_cast<uint32_t>(j < minIdx ? j + 1 : j));
          const int span = iv.end - iv.begin;
          for (int e = 0; e < span; ++e) {
            keys[at] = (static_cast<uint64_t>(ord) << 16) | static_cast<uint64_t>(e);
            vals[at] = (static_cast<uint64_t>(static_cast<uint32_t>(iv.begin + e)) << 32) | (static_cast<uint64_t>(iv.qpos) << 16) | iv.len;
            ++at;
          }
        }
      }
    }
    // ---- stage 2: SA entry -> {transcript, position in the transcript}, one independent load per entry
#pragma unroll 4
    for (uint32_t i = 0; i < total; ++i) {
      const uint64_t v = vals[i];
      const uint2 tp = __ldg(ix.saTidPos + (v >> 32));
      keys[i] |= static_cast<uint64_t>(tp.x) << 32;
      vals[i] = (static_cast<uint64_t>(tp.y) << 32) | (v & 0xffffffffULL);
    }
    // ---- per strand: sort, resolve the transcript segments; hit o of a strand is compacted in place at [b0 + o]:
    //      keys = tid << 32 | pos, vals = chain score (double bits), outMeta = posOff << 16 | nAll << 8 | chain
    uint32_t nOutF = 0, nOutR = 0;
    for (int strand = 0; strand < 2 && total > 0; ++strand) {
      const uint32_t b0 = strand ? totF : 0u, cnt = strand ? totR : totF;
      const int n = strand ? nR : nF;
      if (cnt == 0) continue;
      for (uint32_t i = 1; i < cnt; ++i) {  // insertion sort by (tid, ord, entry); keys are unique
        const uint64_t kx = keys[b0 + i], vx = vals[b0 + i];
        int32_t j = static_cast<int32_t>(i) - 1;
        while (j >= 0 && keys[b0 + j] > kx) { keys[b0 + j + 1] = keys[b0 + j]; vals[b0 + j + 1] = vals[b0 + j]; --j; }
        keys[b0 + j + 1] = kx; vals[b0 + j + 1] = vx;
      }
      int32_t required = n, maxSlack = 0;  // :613-628
      if (n > 1 && o.consensusFraction < 1.0f) {
        const float requiredFrac = __fmul_rn(static_cast<float>(n), o.consensusFraction);
        const int32_t fl = static_cast<int32_t>(floorf(requiredFrac));
        required = fl > 1 ? fl : 1;
        maxSlack = n - required;
      }
      // does any transcript reach the required number of intervals?  (else, with slack, all are kept: :667-686)
      bool keepAll = n == 1;
      if (!keepAll && maxSlack > 0) {
        bool any = false;
        uint32_t i = 0;
        while (i < cnt && !any) {
          const uint32_t tid = static_cast<uint32_t>(keys[b0 + i] >> 32);
          int32_t distinct = 0;
          uint32_t lastOrd = 0xffffffffu;
          while (i < cnt && static_cast<uint32_t>(keys[b0 + i] >> 32) == tid) {
            const uint32_t ord = static_cast<uint32_t>(keys[b0 + i] >> 16) & 0xffffu;
            if (ord != lastOrd) { ++distinct; lastOrd = ord; }
            ++i;
          }
          any = distinct >= required;
        }
        keepAll = !any;
      }
      const bool chain = o.doChaining && n > 1;
      const IntervalRec iv0 = ivs[strand ? nF : 0];
      uint32_t nOut = 0, i = 0;
      while (i < cnt) {
        const uint32_t tid = static_cast<uint32_t>(keys[b0 + i] >> 32);
        uint32_t e = i;
        int32_t distinct = 0;
        uint32_t lastOrd = 0xffffffffu;
        while (e < cnt && static_cast<uint32_t>(keys[b0 + e] >> 32) == tid) {
          const uint32_t ord = static_cast<uint32_t>(keys[b0 + e] >> 16) & 0xffffu;
          if (ord != lastOrd) { ++distinct; lastOrd = ord; }
          ++e;
        }
        if (keepAll || distinct >= required) {
          QARec q;
          q.tid = tid; q.pos = 0; q.nAll = 1; q.chain = 4;
          double score = -1.7976931348623157e308;
          const uint32_t b = b0 + i, en = b0 + e;
          if (chain) {
            q.nAll = chainSegment(w, b, en, readLen, static_cast<int32_t>(readLen), o.considerMultiPos, q, score);
          } else if (n == 1) {  // collectFromSingleInterval (:716-807)
            q.chain = (iv0.len == readLen) ? 0 : 4;
            const int32_t qp = static_cast<int32_t>(iv0.qpos);
            if (o.considerMultiPos) {
              int32_t* out = w.posTmp + b;
              uint32_t c2 = 0;
              for (uint32_t t = b; t < en; ++t) {
                const int32_t hp = static_cast<int32_t>(vPos(vals[t])) - qp;
                int32_t j = static_cast<int32_t>(c2) - 1;
                while (j >= 0 && out[j] > hp) { out[j + 1] = out[j]; --j; }
                out[j + 1] = hp;
                ++c2;
              }
              q.pos = out[0]; q.nAll = c2;
            } else {
              int32_t best = 0x7fffffff;
              for (uint32_t t = b; t < en; ++t) { const int32_t hp = static_cast<int32_t>(vPos(vals[t])) - qp; best = hp < best ? hp : best; }
              q.pos = best;
              if (needPos) w.posTmp[b] = best;
            }
          } else {  // leftmost anchor (:308-322)
            uint32_t bestPos = 0xffffffffu;
            int32_t bestHit = 0;
            for (uint32_t t = b; t < en; ++t) {
              const uint64_t v = vals[t];
              if (vPos(v) < bestPos) { bestPos = vPos(v); bestHit = static_cast<int32_t>(vPos(v) - vQpos(v)); }
            }
            q.pos = bestHit;
            if (needPos) w.posTmp[b] = bestHit;
          }
          keys[b0 + nOut] = (static_cast<uint64_t>(tid) << 32) | static_cast<uint32_t>(q.pos);
          vals[b0 + nOut] = static_cast<uint64_t>(__double_as_longlong(score));
          outMeta[b0 + nOut] = (b << 16) | (q.nAll << 8) | q.chain;
          ++nOut;
        }
        i = e;
      }
      if (strand) nOutR = nOut; else nOutF = nOut;
    }
    // ---- merge by tid (:834-881): equal tid => higher chain score first (forward on a tie); the winner takes the
    //      loser's positions as oppositeStrandPositions.  First pass counts hits and positions.
    uint32_t nFinal = 0, nPosTot = 0;
    if (total > 0) {
      uint32_t a = 0, b = 0;
      while (a < nOutF || b < nOutR) {
        bool takeA, both = false;
        if (b >= nOutR) takeA = true;
        else if (a >= nOutF) takeA = false;
        else {
          const uint32_t ta = static_cast<uint32_t>(keys[a] >> 32), tb = static_cast<uint32_t>(keys[totF + b] >> 32);
          if (ta == tb) { both = true; takeA = true; } else takeA = ta < tb;
        }
        if (both) { nPosTot += ((outMeta[a] >> 8) & 0xffu) + ((outMeta[totF + b] >> 8) & 0xffu); ++a; ++b; }
        else if (takeA) { nPosTot += (outMeta[a] >> 8) & 0xffu; ++a; }
        else { nPosTot += (outMeta[totF + b] >> 8) & 0xffu; ++b; }
        ++nFinal;
      }
    }
    if (!needPos) nPosTot = 0;
    const unsigned pm = __ballot_sync(0xffffffffu, nFinal > 0);
    uint32_t off = 0, poolOff = 0;
    if (pm) {
      int incl = static_cast<int>(nFinal), inclP = static_cast<int>(nPosTot);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d), vp = __shfl_up_sync(0xffffffffu, inclP, d);
        if (lane >= d) { incl += v; inclP += vp; }
      }
      const uint32_t tot = static_cast<uint32_t>(__shfl_sync(0xffffffffu, incl, 31)), totP = static_cast<uint32_t>(__shfl_sync(0xffffffffu, inclP, 31));
      uint32_t wbase = 0, pbase = 0;
      if (lane == 0) { wbase = atomicAdd(P.qaCursor, tot); if (totP) pbase = atomicAdd(P.posCursor, totP); }
      wbase = __shfl_sync(0xffffffffu, wbase, 0); pbase = __shfl_sync(0xffffffffu, pbase, 0);
      off = wbase + static_cast<uint32_t>(incl) - nFinal;
      poolOff = pbase + static_cast<uint32_t>(inclP) - nPosTot;
      bool ok = true;
      if (static_cast<uint64_t>(wbase) + tot > P.qaCap) { if (lane == 0) atomicOr(P.status, kStatQAArenaFull); ok = false; }
      if (static_cast<uint64_t>(pbase) + totP > P.posCap) { if (lane == 0) atomicOr(P.status, kStatPosPoolFull); ok = false; }
      if (!ok) nFinal = 0;
    }
    if (nFinal > 0) {
      uint32_t a = 0, b = 0, wr = 0, run = poolOff;
      while (a < nOutF || b < nOutR) {
        bool takeA, both = false;
        if (b >= nOutR) takeA = true;
        else if (a >= nOutF) takeA = false;
        else {
          const uint32_t ta = static_cast<uint32_t>(keys[a] >> 32), tb = static_cast<uint32_t>(keys[totF + b] >> 32);
          if (ta == tb) { both = true; takeA = !(__longlong_as_double(static_cast<long long>(vals[totF + b])) > __longlong_as_double(static_cast<long long>(vals[a]))); }
          else takeA = ta < tb;
        }
        const uint32_t win = takeA ? a : totF + b, lose = takeA ? totF + b : a;
        const uint64_t kx = keys[win];
        const uint32_t mw = outMeta[win];
        QARec q;
        q.tid = static_cast<uint32_t>(kx >> 32);
        q.pos = static_cast<int32_t>(static_cast<uint32_t>(kx));
        q.nAll = (mw >> 8) & 0xffu; q.chain = static_cast<uint8_t>(mw & 0xffu);
        q.fwd = takeA ? 1 : 0; q.pad = 0;
        q.posOff = 0; q.oppOff = 0; q.nOpp = 0;
        if (both) q.nOpp = (outMeta[lose] >> 8) & 0xffu;
        if (needPos) {
          const int32_t* src = w.posTmp + (mw >> 16);
          for (uint32_t t = 0; t < q.nAll; ++t) P.posPool[run + t] = src[t];
          q.posOff = run;
          run += q.nAll;
          if (both) {
            const int32_t* src2 = w.posTmp + (outMeta[lose] >> 16);
            for (uint32_t t = 0; t < q.nOpp; ++t) P.posPool[run + t] = src2[t];
          }
          q.oppOff = run;
          run += q.nOpp;
        }
        P.qaArena[off + wr] = q;
        ++wr;
        if (both) { ++a; ++b; } else if (takeA) ++a; else ++b;
      }
    }
    if (valid) {
      QASummary out;
      out.qaOff = off; out.nQA = todo ? kTodoMark : nFinal;
      P.qsumm[r] = out;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Size classes.  Under -s / -f a read's intervals expand to ~20 SA entries on average and to 64 and more for a few per cent
// (profiles/r02e): a lane kernel with one fixed strip either leaves half of the reads to the warp-per-read kernel (strip of
// 16: 9.7 of 11.7 ms were spent there) or pays the big strip and the longest read of every warp for all reads.  So the
// reads are grouped by their entry count (counting sort: histogram, then scatter) and the lane kernel is launched per
// size range with a strip and block size to match; consecutive reads of a warp then have (nearly) the same amount of work.
__device__ __forceinline__ int k2Bucket(const MapParams& P, uint64_t r, int maxIv) {
  const ReadSummary s = P.summ[r];
  const int nIv = s.nFwd + s.nRc;
  if (nIv == 0) return 0;
  if (nIv > maxIv) return kK2MaxEntries + 1;
  const IntervalRec* ivs = P.arena + s.ivOff;
  uint32_t tot = 0;
  for (int j = 0; j < nIv; ++j) tot += static_cast<uint32_t>(ivs[j].end - ivs[j].begin);
  if (tot == 0) return 1;
  return tot <= static_cast<uint32_t>(kK2MaxEntries) ? static_cast<int>(tot) : kK2MaxEntries + 1;
}

__global__ void __launch_bounds__(256) k2_class_hist_kernel(MapParams P, int maxIv) {
  __shared__ uint32_t h[kK2Buckets];
  for (int i = threadIdx.x; i < kK2Buckets; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < P.numReads; r += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    atomicAdd(&h[k2Bucket(P, r, maxIv)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < kK2Buckets; i += blockDim.x)
    if (h[i]) atomicAdd(P.classHist + i, h[i]);
}

__global__ void __launch_bounds__(256) k2_class_scatter_kernel(MapParams P, int maxIv, uint32_t* order) {
  __shared__ uint32_t cnt[kK2Buckets], base[kK2Buckets];
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t rounds = (P.numReads + stride - 1) / stride;
  for (uint64_t it = 0; it < rounds; ++it) {
    const uint64_t r = it * stride + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (int i = threadIdx.x; i < kK2Buckets; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    int c = 0;
    uint32_t rank = 0;
    if (r < P.numReads) {
      c = k2Bucket(P, r, maxIv);
      rank = atomicAdd(&cnt[c], 1u);
      if (c == 0) { QASummary z; z.qaOff = 0; z.nQA = 0; P.qsumm[r] = z; }  // nothing to resolve
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kK2Buckets; i += blockDim.x) {
      uint32_t off = 0;  // exclusive scan of the histogram
      for (int j = 0; j < i; ++j) off += P.classHist[j];
      base[i] = off + (cnt[i] ? atomicAdd(P.classHist + kK2Buckets + i, cnt[i]) : 0u);
    }
    __syncthreads();
    if (r < P.numReads) order[base[c] + rank] = static_cast<uint32_t>(r);
    __syncthreads();
  }
}

} // namespace rapmap_b200