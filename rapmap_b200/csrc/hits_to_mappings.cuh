// Kernel 2 — hit resolution for one read per warp.
//
// Replaces hit_manager::hitsToMappingsSimple (reference src/HitManager.cpp:691-882): the single-interval
// expansion (:716-807), intersectSAHits + intersectSAIntervalWithOutput (:587-689, :449-493),
// collectHitsSimpleSA (:84-326, leftmost anchor or chain DP) and the fwd/rc merge (:834-881).
//
// B200 mapping: the reference walks SA entries one by one into a std::map<tid, ProcessedSAHit>.  Here the
// warp expands all SA entries of a strand at once (lane = entry: SA[i] -> rank record -> txpOffsets, three
// dependent loads per lane, 32 in flight), sorts the (tid, interval order, entry) keys with a warp bitonic
// network (shared memory for the common <= 64 entries, an L2-resident global scratch strip otherwise) and
// resolves each transcript segment with one lane.  Map iteration order == ascending tid == sorted order.
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
  // per-warp work strip (global): keys u64[cap] | vals u64[cap] | segs u32[cap+1] | qa QARec[cap] | dbl double[cap] ...
  uint8_t* scratch;
  uint32_t scratchEntries;   // entries per warp in the global strip
  uint64_t scratchStride;    // bytes per warp
  uint32_t smemEntries;      // entries per warp in shared memory
  uint32_t* status;
};

// Work-area view (either shared memory or the warp's global strip).
struct WorkArea {
  uint64_t* keys;   // tid << 32 | ord << 16 | entry
  uint64_t* vals;   // pos << 32 | qpos << 16 | len
  uint32_t* segs;   // segment starts
  QARec* qa;        // resolved hits of this read (fwd list then rc list)
  uint32_t cap;
};

__host__ __device__ inline uint64_t workAreaBytes(uint32_t entries) {
  // keys + vals + segs(+1) + qa, 16-byte aligned
  uint64_t b = static_cast<uint64_t>(entries) * (8 + 8 + 4 + sizeof(QARec)) + 16;
  return (b + 15) / 16 * 16;
}

__device__ __forceinline__ WorkArea carve(uint8_t* p, uint32_t entries) {
  WorkArea w;
  w.keys = reinterpret_cast<uint64_t*>(p);
  w.vals = w.keys + entries;
  w.qa = reinterpret_cast<QARec*>(w.vals + entries);
  w.segs = reinterpret_cast<uint32_t*>(w.qa + entries);
  w.cap = entries;
  return w;
}

__device__ __forceinline__ void warpBitonicSort(uint64_t* keys, uint64_t* vals, int n2, int lane) {
  for (int kk = 2; kk <= n2; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n2 >> 1); t += 32) {
        int i = ((t / j) * (j << 1)) + (t % j);
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

// Resolves one strand: appends QARecs (ascending tid) to w.qa[nOut...].  Returns false on work-area overflow.
// No chaining (doChaining == false): leftmost anchor per active transcript (HitManager.cpp:308-322) or per
// transcript of the single interval (:716-807).
__device__ __forceinline__ bool resolveStrand(const MapParams& P, WorkArea& w, const IntervalRec* ivs, int nIv, bool isFw, uint32_t readLen,
                                              uint8_t mateStatus, int lane, uint32_t& nOut, bool& overflow) {
  const DeviceIndex& ix = P.ix;
  // ---- order of processing: smallest-span interval first (first wins on ties, :636-641), others in original order
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
  if (total + nOut > w.cap) { overflow = true; return false; }
  // ---- expand: lane = SA entry
  uint32_t base = 0;
  for (int j = 0; j < nIv; ++j) {
    IntervalRec iv = ivs[j];
    uint32_t ord = (nIv == 1) ? 0u : (j == minIdx ? 0u : static_cast<uint32_t>(j < minIdx ? j + 1 : j));
    int span = iv.end - iv.begin;
    for (int e = lane; e < span; e += 32) {
      int32_t g = __ldg(ix.SA + iv.begin + e);
      uint32_t tid = transcriptAt(ix, g);
      int32_t pos = g - __ldg(ix.txpOffsets + tid);
      w.keys[base + e] = (static_cast<uint64_t>(tid) << 32) | (static_cast<uint64_t>(ord) << 16) | static_cast<uint64_t>(e);
      w.vals[base + e] = (static_cast<uint64_t>(static_cast<uint32_t>(pos)) << 32) | (static_cast<uint64_t>(iv.qpos) << 16) | iv.len;
    }
    base += static_cast<uint32_t>(span);
  }
  __syncwarp();
  // ---- sort by (tid, ord, entry)
  if (total > 1) {
    int n2 = 1;
    while (n2 < static_cast<int>(total)) n2 <<= 1;
    if (static_cast<uint32_t>(n2) > w.cap) { overflow = true; return false; }
    for (int i = total + lane; i < n2; i += 32) { w.keys[i] = ~0ULL; w.vals[i] = 0; }
    __syncwarp();
    warpBitonicSort(w.keys, w.vals, n2, lane);
  }
  // ---- segment heads
  uint32_t nSeg = 0;
  for (uint32_t b0 = 0; b0 < total; b0 += 32) {
    uint32_t i = b0 + lane;
    bool head = false;
    if (i < total) head = (i == 0) || ((w.keys[i] >> 32) != (w.keys[i - 1] >> 32));
    unsigned m = __ballot_sync(0xffffffffu, head);
    if (head) w.segs[nSeg + __popc(m & ((1u << lane) - 1u))] = i;
    nSeg += __popc(m);
  }
  if (lane == 0) w.segs[nSeg] = total;
  __syncwarp();
  // ---- one lane per transcript segment
  const uint32_t required = static_cast<uint32_t>(nIv);  // strict intersection: present in every interval (:613-628 with consensusFraction == 1)
  for (uint32_t s0 = 0; s0 < nSeg; s0 += 32) {
    uint32_t s = s0 + lane;
    bool active = false;
    QARec q;
    if (s < nSeg) {
      uint32_t b = w.segs[s], e = w.segs[s + 1];
      uint32_t tid = static_cast<uint32_t>(w.keys[b] >> 32);
      uint32_t distinct = 0, lastOrd = 0xffffffffu;
      uint32_t bestPos = 0xffffffffu;
      int32_t bestHit = 0;
      for (uint32_t i = b; i < e; ++i) {
        uint32_t ord = static_cast<uint32_t>(w.keys[i] >> 16) & 0xffffu;
        if (ord != lastOrd) { ++distinct; lastOrd = ord; }
        uint64_t v = w.vals[i];
        uint32_t pos = static_cast<uint32_t>(v >> 32);
        if (pos < bestPos) {  // std::min_element keeps the first minimum in tqvec order == sorted (ord, entry) order
          bestPos = pos;
          bestHit = static_cast<int32_t>(pos) - static_cast<int32_t>((v >> 16) & 0xffffu);
        }
      }
      active = (nIv == 1) || (distinct >= required);
      q.tid = tid; q.pos = bestHit; q.posOff = 0; q.nAll = 1; q.nOpp = 0; q.fwd = isFw ? 1 : 0; q.pad = 0;
      uint8_t cs = 4;  // REGULAR
      if (nIv == 1) cs = (ivs[0].len == readLen) ? 0 : 4;  // PERFECT iff the single MMP spans the read (:741-746)
      q.chain = cs;
    }
    unsigned m = __ballot_sync(0xffffffffu, active);
    if (active) w.qa[nOut + __popc(m & ((1u << lane) - 1u))] = q;
    nOut += __popc(m);
  }
  __syncwarp();
  return true;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) hits_to_mappings_kernel(MapParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint64_t gw = static_cast<uint64_t>(blockIdx.x) * WARPS + warp;
  uint8_t* smemBase = smem + static_cast<size_t>(warp) * workAreaBytes(P.smemEntries);
  uint8_t* globBase = P.scratch + gw * P.scratchStride;

  for (uint64_t r = gw; r < P.numReads; r += static_cast<uint64_t>(gridDim.x) * WARPS) {
    ReadSummary s = P.summ[r];
    QASummary out;
    out.qaOff = 0; out.nQA = 0;
    const int nF = s.nFwd, nR = s.nRc;
    if (nF + nR == 0) {
      if (lane == 0) P.qsumm[r] = out;
      continue;
    }
    const IntervalRec* ivs = P.arena + s.ivOff;
    uint32_t total = 0;
    for (int j = 0; j < nF + nR; ++j) total += static_cast<uint32_t>(ivs[j].end - ivs[j].begin);
    // pow2 padding of the larger strand must fit as well
    uint32_t need = 1;
    while (need < total) need <<= 1;
    need = need > total ? need : total;
    WorkArea w = (need <= P.smemEntries) ? carve(smemBase, P.smemEntries) : carve(globBase, P.scratchEntries);
    const uint8_t mateStatus = P.pairedInput ? (r >= P.numPairs ? 2 : 1) : 0;
    uint32_t nOut = 0;
    bool overflow = false;
    uint32_t nFwdOut = 0;
    if (nF > 0) resolveStrand(P, w, ivs, nF, true, s.readLen, mateStatus, lane, nOut, overflow);
    nFwdOut = nOut;
    if (!overflow && nR > 0) resolveStrand(P, w, ivs + nF, nR, false, s.readLen, mateStatus, lane, nOut, overflow);
    if (overflow) {
      if (lane == 0) { atomicOr(P.status, kStatScratchFull); P.qsumm[r] = out; }
      continue;
    }
    uint32_t nFinal = nOut;
    // ---- merge forward and reverse-complement lists (HitManager.cpp:834-881); rare, done by lane 0.
    // Without chain scores every tie keeps the forward hit (stable inplace_merge, equal chainScore).
    if (nFwdOut > 0 && nOut > nFwdOut) {
      if (lane == 0) {
        // merged order written behind the two lists, then moved to the front
        uint32_t a = 0, b = nFwdOut, o = nOut;
        if (2 * nOut <= w.cap) {
          while (a < nFwdOut || b < nOut) {
            if (b >= nOut || (a < nFwdOut && w.qa[a].tid <= w.qa[b].tid)) {
              if (a < nFwdOut && b < nOut && w.qa[a].tid == w.qa[b].tid) ++b;  // drop the rc duplicate
              w.qa[o++] = w.qa[a++];
            } else {
              w.qa[o++] = w.qa[b++];
            }
          }
          nFinal = o - nOut;
          for (uint32_t i = 0; i < nFinal; ++i) w.qa[i] = w.qa[nOut + i];
        } else {
          nFinal = 0xffffffffu;
        }
      }
      nFinal = __shfl_sync(0xffffffffu, nFinal, 0);
      if (nFinal == 0xffffffffu) {
        if (lane == 0) { atomicOr(P.status, kStatScratchFull); P.qsumm[r] = out; }
        continue;
      }
      __syncwarp();
    }
    // ---- publish
    uint32_t off = 0;
    if (nFinal > 0) {
      if (lane == 0) off = atomicAdd(P.qaCursor, nFinal);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off + nFinal > P.qaCap) {
        if (lane == 0) atomicOr(P.status, kStatQAArenaFull);
        nFinal = 0;
      } else {
        for (uint32_t i = lane; i < nFinal; i += 32) P.qaArena[off + i] = w.qa[i];
      }
    }
    if (lane == 0) { out.qaOff = off; out.nQA = nFinal; P.qsumm[r] = out; }
    __syncwarp();
  }
}

} // namespace rapmap_b200
