// Kernel 1, regrouped form — the SA-lookup kernel with the reads of a block sorted by what they need next.
//
// Same walk, same helpers and same results as sa_collect_lane_kernel (sa_collect_lane.cuh), different scheduling.
// There a lane keeps its read from start to finish and every loop trip runs the lookup phase for the ~19 lanes that
// need it and the probe phase for the ~7 that need it: 7.3 of 32 threads active per instruction (profiles/r01h) and a
// trip as long as both phases together.  Here the state of a read lives in shared memory (17 words + its packed
// bases), a block owns SLOTS reads, and two queues hold the slots whose next expensive step is a LOOKUP (k-mer +
// reverse complement: filter, then table) or a PROBE (one suffix comparison of extendSearchNaive).  Per round the
// warps take 32 slots of one kind at a time, load their state into registers, run that one step plus the cheap
// bookkeeping that follows it (interval record, strand sequencing, publish, refill from the batch), store the state and
// push the slot on the queue of its next step.  Every expensive instruction is issued for a full warp of reads that
// want exactly it.
#pragma once
#include "sa_collect_lane.cuh"

namespace rapmap_b200 {

static constexpr int kRgStateWords = 17;

__host__ __device__ inline uint32_t regroupSmemBytes(uint32_t nw, uint32_t slots) {
  return slots * (16u * nw + 4u * kRgStateWords + 3u * 2u * 2u) + 64u;
}

struct RgState {
  uint32_t r, flags, fwdHit, rcHit, fwdCov, rcCov;
  int st, L, rb, lbIn, ubIn, l, rr, lcpLP, lcpRP, prevILow, prevIHigh, mlen, b0, b1, mQ, pass, guard, prevMMPEnd, nF, nR;
};

template <int SLOTS>
__device__ __forceinline__ void rgLoad(const uint32_t* stW, int s, RgState& S) {
  const uint32_t* p = stW + s;
  S.r = p[0];
  uint32_t w = p[1 * SLOTS]; S.L = static_cast<int>(w & 0xffffu); S.rb = static_cast<int>(w >> 16);
  w = p[2 * SLOTS]; S.flags = w & 0xffffu; S.st = static_cast<int>((w >> 16) & 0xfu); S.pass = static_cast<int>((w >> 20) & 0x3u); S.guard = static_cast<int>(w >> 22);
  S.lbIn = static_cast<int>(p[3 * SLOTS]); S.ubIn = static_cast<int>(p[4 * SLOTS]); S.l = static_cast<int>(p[5 * SLOTS]); S.rr = static_cast<int>(p[6 * SLOTS]);
  S.b0 = static_cast<int>(p[7 * SLOTS]); S.b1 = static_cast<int>(p[8 * SLOTS]);
  w = p[9 * SLOTS]; S.lcpLP = static_cast<int>(w & 0xffffu); S.lcpRP = static_cast<int>(w >> 16);
  w = p[10 * SLOTS]; S.prevILow = static_cast<int>(w & 0xffffu); S.prevIHigh = static_cast<int>(w >> 16);
  w = p[11 * SLOTS]; S.mlen = static_cast<int>(w & 0xffffu); S.mQ = static_cast<int>(w >> 16);
  w = p[12 * SLOTS]; S.prevMMPEnd = static_cast<int>(w & 0xffffu); S.nF = static_cast<int>(w >> 16);
  w = p[13 * SLOTS]; S.nR = static_cast<int>(w & 0xffffu); S.fwdHit = w >> 16;
  S.rcHit = p[14 * SLOTS];
  S.fwdCov = p[15 * SLOTS]; S.rcCov = p[16 * SLOTS];
}

template <int SLOTS>
__device__ __forceinline__ void rgStore(uint32_t* stW, int s, const RgState& S) {
  uint32_t* p = stW + s;
  p[0] = S.r;
  p[1 * SLOTS] = static_cast<uint32_t>(S.L) | (static_cast<uint32_t>(S.rb) << 16);
  p[2 * SLOTS] = (S.flags & 0xffffu) | (static_cast<uint32_t>(S.st) << 16) | (static_cast<uint32_t>(S.pass) << 20) | (static_cast<uint32_t>(S.guard) << 22);
  p[3 * SLOTS] = static_cast<uint32_t>(S.lbIn); p[4 * SLOTS] = static_cast<uint32_t>(S.ubIn); p[5 * SLOTS] = static_cast<uint32_t>(S.l); p[6 * SLOTS] = static_cast<uint32_t>(S.rr);
  p[7 * SLOTS] = static_cast<uint32_t>(S.b0); p[8 * SLOTS] = static_cast<uint32_t>(S.b1);
  p[9 * SLOTS] = static_cast<uint32_t>(S.lcpLP) | (static_cast<uint32_t>(S.lcpRP) << 16);
  p[10 * SLOTS] = static_cast<uint32_t>(S.prevILow) | (static_cast<uint32_t>(S.prevIHigh) << 16);
  p[11 * SLOTS] = static_cast<uint32_t>(S.mlen) | (static_cast<uint32_t>(S.mQ) << 16);
  p[12 * SLOTS] = static_cast<uint32_t>(S.prevMMPEnd) | (static_cast<uint32_t>(S.nF) << 16);
  p[13 * SLOTS] = static_cast<uint32_t>(S.nR) | (S.fwdHit << 16);
  p[14 * SLOTS] = S.rcHit;
  p[15 * SLOTS] = S.fwdCov; p[16 * SLOTS] = S.rcCov;
}

#ifndef RAPMAP_RG_MINB
#define RAPMAP_RG_MINB 3
#endif

template <int NT, int SLOTS>
__global__ void __launch_bounds__(NT, RAPMAP_RG_MINB) sa_collect_regroup_kernel(LaneParams P) {
  extern __shared__ __align__(16) uint8_t smemRg[];
  const int nw = static_cast<int>(P.nw);
  uint4* packW = reinterpret_cast<uint4*>(smemRg);                                   // [nw][SLOTS]
  uint32_t* stW = reinterpret_cast<uint32_t*>(packW + static_cast<size_t>(nw) * SLOTS);  // [17][SLOTS]
  uint16_t* qBuf = reinterpret_cast<uint16_t*>(stW + kRgStateWords * SLOTS);        // [3 queue sets][2 kinds][SLOTS]
  uint32_t* ctl = reinterpret_cast<uint32_t*>(qBuf + 6 * SLOTS);                    // [0..5] queue counts, [6..8] chunk counters
  const int lane = threadIdx.x & 31;
  const int k = static_cast<int>(P.ix.k);
  const DevOpts& o = P.opts;
  const bool useCov = o.disableNIP && o.strictCheck;  // include/SACollector.hpp:138
  const bool voteMode = o.strictCheck && !useCov;
  const size_t slot0 = static_cast<size_t>(blockIdx.x) * SLOTS;

  uint32_t chunkBase = 0, chunkLeft = 0;  // warp-uniform: reserved slice of the interval arena

  // ---- converged tail of every step: publish finished reads, give their slots new reads, store, enqueue
  auto finishStep = [&](RgState& S, int s, bool have, int nxt) {
    IntervalRec* scr = P.ivScratch + (slot0 + static_cast<size_t>(s)) * 2 * P.ivStride;
    uint32_t* votes = voteMode ? P.voteScratch + (slot0 + static_cast<size_t>(s)) * 3 * P.voteWords : nullptr;
    // publish (same rules as sa_collect_lane_kernel)
    {
      const bool fin = have && S.st == LST_FINAL;
      int tot = 0;
      if (fin) {
        if (S.flags & LF_FOUND) {
          if (useCov) {  // strand decision by coverage (:283-288)
            if (S.fwdCov > S.rcCov + static_cast<uint32_t>(o.strictCheckSlack)) S.nR = 0;
            else if (S.rcCov > S.fwdCov + static_cast<uint32_t>(o.strictCheckSlack)) S.nF = 0;
          } else if (o.strictCheck) {  // k-mer "spot check" vote (:289-337)
            if (S.fwdHit > 0u && S.rcHit == 0u) S.nR = 0;
            else if (S.rcHit > 0u && S.fwdHit == 0u) S.nF = 0;
            else {
              int fs = 0, rs = 0;
              for (uint32_t j = 0; j < P.voteWords; ++j) {
                const int tested = __popc(votes[j]);
                fs += 2 * __popc(votes[P.voteWords + j]) - tested;
                rs += 2 * __popc(votes[2 * P.voteWords + j]) - tested;
              }
              if (fs > rs) S.nR = 0;
              else if (rs > fs) S.nF = 0;
            }
          }
          if (o.covReq > 0.0 && o.disableNIP) {  // :343-358
            if (S.nF > 0 && (static_cast<double>(S.fwdCov) / static_cast<double>(S.L)) < o.covReq) S.nF = 0;
            if (S.nR > 0 && (static_cast<double>(S.rcCov) / static_cast<double>(S.L)) < o.covReq) S.nR = 0;
          }
        } else { S.nF = 0; S.nR = 0; }
        tot = S.nF + S.nR;
      }
      const unsigned pm = __ballot_sync(0xffffffffu, fin && tot > 0);
      uint32_t off = 0;
      if (pm) {
        int incl = fin ? tot : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        const uint32_t total = static_cast<uint32_t>(__shfl_sync(0xffffffffu, incl, 31));
        if (total > chunkLeft) {
          const uint32_t grab = total > RAPMAP_LANE_CHUNK ? total : RAPMAP_LANE_CHUNK;
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(P.arenaCursor, grab);
          chunkBase = __shfl_sync(0xffffffffu, base, 0);
          chunkLeft = grab;
        }
        off = chunkBase + static_cast<uint32_t>(incl - tot);
        chunkBase += total;
        chunkLeft -= total;
      }
      if (fin) {
        if (tot > 0) {
          if (S.flags & LF_OVF) atomicOr(P.status, kStatIvScratchFull);
          else if (static_cast<uint64_t>(off) + static_cast<uint64_t>(tot) > P.arenaCap) atomicOr(P.status, kStatIntervalArenaFull);
          else {
            for (int i = 0; i < S.nF; ++i) P.arena[off + i] = scr[i];
            for (int i = 0; i < S.nR; ++i) P.arena[off + S.nF + i] = scr[P.ivStride + i];
          }
        }
        ReadSummary sm;
        sm.ivOff = off; sm.nFwd = static_cast<uint16_t>(S.nF); sm.nRc = static_cast<uint16_t>(S.nR);
        sm.readLen = static_cast<uint16_t>(S.L); sm.found = (S.flags & LF_FOUND) ? 1 : 0; sm.pad = 0;
        P.summ[S.r] = sm;
        S.st = LST_IDLE;
      }
    }
    // refill: idle slots take the next reads of the batch
    for (;;) {
      const unsigned idle = __ballot_sync(0xffffffffu, have && S.st == LST_IDLE);
      if (idle == 0u) break;
      const int leader = __ffs(idle) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(P.readCursor, static_cast<uint32_t>(__popc(idle)));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (have && S.st == LST_IDLE) {
        const uint64_t nr = static_cast<uint64_t>(base) + __popc(idle & ((1u << lane) - 1u));
        if (nr >= P.reads.numReads) S.st = LST_EXIT;
        else {
          S.r = static_cast<uint32_t>(nr);
          const int mate = nr >= P.reads.n ? 1 : 0;
          const uint64_t ri = nr - static_cast<uint64_t>(mate) * P.reads.n;
          uint32_t len = P.reads.fixedLen;
          if (P.reads.off[mate]) len = static_cast<uint32_t>(P.reads.off[mate][ri + 1] - P.reads.off[mate][ri]);
          if (len > P.maxReadLen) {
            atomicOr(P.status, kStatReadTooLong);
            ReadSummary sm;
            sm.ivOff = 0; sm.nFwd = 0; sm.nRc = 0; sm.readLen = 0; sm.found = 0; sm.pad = 0;
            P.summ[S.r] = sm;
          } else {
            const uint4* src = P.packed + static_cast<size_t>(nr) * P.nw;
            for (int j = 0; j < nw; ++j) packW[static_cast<size_t>(j) * SLOTS + s] = __ldg(src + j);
            if (voteMode) for (uint32_t j = 0; j < 3 * P.voteWords; ++j) votes[j] = 0u;
            S.L = static_cast<int>(len);
            S.rb = 0; S.flags = 0; S.nF = 0; S.nR = 0; S.fwdHit = 0; S.rcHit = 0; S.fwdCov = 0; S.rcCov = 0;
            S.lbIn = 0; S.ubIn = 0; S.l = 0; S.rr = 0; S.lcpLP = 0; S.lcpRP = 0; S.prevILow = 0; S.prevIHigh = 0; S.mlen = 0;
            S.b0 = 0; S.b1 = 0; S.mQ = 0; S.pass = 0; S.guard = 0; S.prevMMPEnd = 0;
            S.st = LST_SCAN;
          }
        }
      }
    }
    // store + enqueue by the next expensive step
    if (have) rgStore<SLOTS>(stW, s, S);
    const bool toD = have && S.st == LST_EXT;
    const bool toB = have && (S.st == LST_SCAN || S.st == LST_WSTART || S.st == LST_MM);
    const unsigned mB = __ballot_sync(0xffffffffu, toB), mD = __ballot_sync(0xffffffffu, toD);
    uint32_t bB = 0, bD = 0;
    if (lane == 0) {
      if (mB) bB = atomicAdd(&ctl[nxt * 2 + 0], static_cast<uint32_t>(__popc(mB)));
      if (mD) bD = atomicAdd(&ctl[nxt * 2 + 1], static_cast<uint32_t>(__popc(mD)));
    }
    bB = __shfl_sync(0xffffffffu, bB, 0); bD = __shfl_sync(0xffffffffu, bD, 0);
    const unsigned lt = (1u << lane) - 1u;
    if (toB) qBuf[(nxt * 2 + 0) * SLOTS + bB + __popc(mB & lt)] = static_cast<uint16_t>(s);
    if (toD) qBuf[(nxt * 2 + 1) * SLOTS + bD + __popc(mD & lt)] = static_cast<uint16_t>(s);
  };

  // ---- cheap transitions after an expensive step, until the read needs a lookup, a probe, or is finished
  auto closure = [&](RgState& S, int s) {
    IntervalRec* scr = P.ivScratch + (slot0 + static_cast<size_t>(s)) * 2 * P.ivStride;
    for (int it = 0; it < 8; ++it) {
      if (S.st == LST_EXTINIT) {
        S.lbIn = S.lbIn - 1 > 0 ? S.lbIn - 1 : 0;  // :553
        const bool firstAttempt = o.doChaining ? (S.rb == 0) : true;
        const int endPos = firstAttempt ? S.L : min(S.rb + k + o.maxMMPExtension, S.L);
        S.mQ = endPos - S.rb;
        S.flags = (S.flags & ~(LF_FIRST | LF_SECOND)) | (firstAttempt ? LF_FIRST : 0u);
        S.pass = (S.ubIn - S.lbIn == 2) ? 3 : 0;
        S.l = S.lbIn; S.rr = S.ubIn; S.lcpLP = k; S.lcpRP = k; S.prevILow = k; S.prevIHigh = k; S.mlen = k; S.guard = 0;
        S.st = LST_EXT;
      }
      if (S.st == LST_EXTDONE) {
        if (o.doChaining && (S.flags & LF_FIRST) && !(S.flags & LF_SECOND) && !(S.mlen >= S.L) && S.mlen >= k + o.maxMMPExtension) {  // :568-575
          S.mQ = min(S.rb + k + o.maxMMPExtension, S.L) - S.rb;
          S.flags |= LF_SECOND;
          S.pass = (S.ubIn - S.lbIn == 2) ? 3 : 0;
          S.l = S.lbIn; S.rr = S.ubIn; S.lcpLP = k; S.lcpRP = k; S.prevILow = k; S.prevIHigh = k; S.mlen = k; S.guard = 0;
          S.st = LST_EXT;
        } else {
          S.st = LST_POSTMM;
          const bool rc = (S.flags & LF_RC) != 0u;
          if (S.b1 > S.b0 && (S.b1 - S.b0) < o.maxInterval) {  // :578
            const int idx = rc ? S.nR : S.nF;
            if (static_cast<uint32_t>(idx) < P.ivStride) {
              IntervalRec rec;
              rec.begin = S.b0; rec.end = S.b1; rec.len = static_cast<uint16_t>(S.mlen); rec.qpos = static_cast<uint16_t>(S.rb);
              scr[(rc ? P.ivStride : 0u) + idx] = rec;
            } else S.flags |= LF_OVF;
            if (rc) ++S.nR; else ++S.nF;
            const int correction = S.prevMMPEnd > S.rb ? S.prevMMPEnd - S.rb : 0;
            if (rc) S.rcCov += static_cast<uint32_t>(S.mlen - correction); else S.fwdCov += static_cast<uint32_t>(S.mlen - correction);
            S.prevMMPEnd = S.rb + S.mlen;
            if (S.rb + S.mlen < S.L) S.st = LST_MM;
          }
          S.lbIn = S.b0; S.ubIn = S.b1;
        }
      }
      if (S.st == LST_POSTMM) {
        if ((S.flags & LF_LAST) || S.rb + S.mlen >= S.L) S.st = LST_WALKEND;  // :623, :630
        else {  // :634-657 next start: MMP skip, or the NIP skip when --noSensitive
          const int mismatchPos = S.rb + S.mlen;
          const int lce = o.disableNIP ? S.mlen : lceLane(P.ix.SA, P.ix.text, P.ix.n, S.lbIn, static_cast<int64_t>(S.ubIn) - 1, S.mlen, S.L - mismatchPos);
          const int skipMatch = mismatchPos - (k - 1), skipLCE = S.rb + lce - (k - 1);
          S.rb = skipMatch > skipLCE ? skipMatch : skipLCE;
          if (!o.disableNIP && lce > S.mlen && S.L > k) S.rb = S.rb < S.L - k ? S.rb : S.L - k;
          if (S.rb + k == S.L) S.flags |= LF_LAST;  // :663
          S.st = LST_WSTART;
        }
      }
      if (S.st == LST_WALKEND) {  // strand sequencing of SACollector::operator(), :247-281
        uint32_t stage = (S.flags & LF_STAGE_MASK) >> LF_STAGE_SHIFT;
        S.st = LST_FINAL;
        if (stage == 0u) {
          stage = 1u;
          if (S.fwdHit) {
            S.flags = (S.flags | LF_DIDFWD) & ~(LF_RC | LF_LAST);
            S.prevMMPEnd = 0;
            S.st = LST_EXTINIT;
          }
        }
        if (S.st == LST_FINAL && stage == 1u) {
          stage = 2u;
          const bool checkRC = useCov ? (S.rcHit > 0u) : (S.rcHit >= S.fwdHit);  // :256
          if (checkRC) { S.flags = (S.flags | LF_RC) & ~LF_LAST; S.rb = 0; S.prevMMPEnd = 0; S.st = LST_WSTART; }
        }
        if (S.st == LST_FINAL && stage == 2u) {
          stage = 3u;
          const bool checkFwd = useCov ? (S.fwdHit > 0u) : (S.fwdHit >= S.rcHit);  // :270
          if (!(S.flags & LF_DIDFWD) && checkFwd) { S.flags &= ~(LF_RC | LF_LAST); S.rb = 0; S.prevMMPEnd = 0; S.st = LST_WSTART; }
        }
        S.flags = (S.flags & ~LF_STAGE_MASK) | (stage << LF_STAGE_SHIFT);
      }
      if (S.st != LST_EXTINIT && S.st != LST_EXTDONE && S.st != LST_POSTMM && S.st != LST_WALKEND) break;
    }
  };

  // ---- LOOKUP step: advance to the next k-mer that needs the table (filter-proven double misses are consumed), then the table
  auto lookupStep = [&](RgState& S, int s, bool have) {
    const uint4* sm = packW + s;
    uint32_t* votes = voteMode ? P.voteScratch + (slot0 + static_cast<size_t>(s)) * 3 * P.voteWords : nullptr;
    bool ready = false, knownM = false, knownC = false;
    uint64_t w = 0;
    int lookPos = 0;
    for (int spin = 0; spin < RAPMAP_LANE_SPIN; ++spin) {
      if (!have || ready || !(S.st == LST_SCAN || S.st == LST_WSTART || S.st == LST_MM)) break;
      const bool rc = (S.flags & LF_RC) != 0u;
      for (;;) {
        if (S.st == LST_MM) lookPos = S.rb + S.mlen - (k - 1);
        else {
          if (S.rb + k > S.L) { S.st = (S.st == LST_SCAN) ? LST_FINAL : LST_WALKEND; break; }
          lookPos = S.rb;
          if (S.st == LST_SCAN && !(S.flags & LF_NOMOREN)) {
            const int ip = findNLane<SLOTS>(sm, S.L, false, S.rb);
            if (ip == INT_MAX) S.flags |= LF_NOMOREN;
            else if (ip <= S.rb + k) { S.rb = ip + 1; continue; }
          }
        }
        const bool valid = kmerAt<SLOTS>(sm, nw, S.L, k, rc, lookPos, w);
        if (S.st == LST_MM) {
          if (valid) ready = true; else S.st = LST_POSTMM;
          break;
        }
        if (S.st == LST_WSTART && !valid) {
          const int ip = findNLane<SLOTS>(sm, S.L, rc, S.rb);
          if (ip < S.rb + k) { S.rb = ip + 1; continue; }
        }
        if (isHomopolymer(w, k)) { ++S.rb; continue; }
        ready = true;
        break;
      }
      if (!ready || P.ix.filter == nullptr) break;
      {
        uint64_t wa, wb;
        uint32_t ma, mb;
        filterSlot(mix64(w), P.ix.filterShift, wa, ma);
        filterSlot(mix64(kmerRC(w, k)), P.ix.filterShift, wb, mb);
        const uint32_t fa = ldgKeep(P.ix.filter + wa), fb = ldgKeep(P.ix.filter + wb);
        knownM = (fa & ma) != ma;
        knownC = (fb & mb) != mb;
      }
      if (!(knownM && knownC)) break;
      ready = false; knownM = false; knownC = false;
      if (voteMode && S.st != LST_SCAN) voteLane(votes, P.voteWords, rc, lookPos, S.L, k, false, false);
      if (S.st == LST_MM) S.st = LST_POSTMM; else ++S.rb;
    }
    __syncwarp();
    if (ready) {
      int2 fm, fc;
      hashFind2(P.ix, w, kmerRC(w, k), knownM, knownC, fm, fc);
      const bool hm = fm.x >= 0, hc = fc.x >= 0;
      const bool rc = (S.flags & LF_RC) != 0u;
      if (S.st == LST_SCAN) {
        if (hm) { ++S.fwdHit; if (hc) ++S.rcHit; }
        if (hc && !S.fwdHit) ++S.rcHit;
        if (S.fwdHit + S.rcHit > 0u) {
          if (voteMode) voteLane(votes, P.voteWords, false, lookPos, S.L, k, hm, hc);
          S.flags |= LF_FOUND;
          S.lbIn = fm.x; S.ubIn = fm.y;
          S.st = LST_WALKEND;
        } else ++S.rb;
      } else {
        if (rc) { S.rcHit += hm ? 1u : 0u; S.fwdHit += hc ? 1u : 0u; } else { S.fwdHit += hm ? 1u : 0u; S.rcHit += hc ? 1u : 0u; }
        if (voteMode) voteLane(votes, P.voteWords, rc, lookPos, S.L, k, hm, hc);
        if (S.st == LST_WSTART) {
          if (!hm) ++S.rb;
          else { S.lbIn = fm.x; S.ubIn = fm.y; S.st = LST_EXTINIT; }
        } else S.st = LST_POSTMM;
      }
    }
  };

  // ---- PROBE step: one suffix comparison of the three binary searches (include/SASearcher.hpp:87-309)
  auto probeStep = [&](RgState& S, int s) {
    const uint4* sm = packW + s;
    int cc, i0, m, sentIdx = -1;
    uint32_t sent = 0;
    if (S.pass == 3) { cc = S.lbIn + 1; i0 = k; }
    else { cc = static_cast<int>((static_cast<int64_t>(S.l) + S.rr) >> 1); i0 = S.lcpLP < S.lcpRP ? S.lcpLP : S.lcpRP; }
    if (S.pass == 1 || S.pass == 2) { m = S.mlen + 1; sentIdx = m - 1; sent = S.pass == 1 ? '#' : '{'; }
    else m = S.mQ;
    const int32_t t = __ldg(P.ix.SA + cc);
    int rel;
    const int i = cmpSuffix<SLOTS>(P, sm, S.r, S.L, (S.flags & LF_RC) != 0u, S.rb, m, t, i0, sentIdx, sent, rel);
    if (S.pass == 3) {
      S.b0 = S.lbIn + 1; S.b1 = S.ubIn; S.mlen = i;
      S.st = LST_EXTDONE;
    } else if (S.pass == 0) {
      bool plt = true;
      if (rel < 0) { if (i > S.prevIHigh) S.prevIHigh = i; }
      else if (rel > 0) { if (i > S.prevILow) S.prevILow = i; plt = false; }
      else if (i == m || static_cast<int64_t>(t) + i == P.ix.n) { if (i > S.prevIHigh) S.prevIHigh = i; }
      bool fin = false;
      if (plt) { if (cc == S.l + 1) fin = true; else { S.rr = cc; S.lcpRP = i; } }
      else { if (cc == S.rr - 1) fin = true; else { S.l = cc; S.lcpLP = i; } }
      if (fin) S.mlen = max(max(i, S.prevILow), S.prevIHigh);
      else if (++S.guard >= 80) fin = true;
      if (fin) { S.pass = 1; S.l = S.lbIn; S.rr = S.ubIn; S.lcpLP = k; S.lcpRP = k; S.guard = 0; }
    } else {
      bool fin = false;
      int bnd = S.ubIn;
      if (rel <= 0) { if (cc == S.l + 1) { bnd = cc; fin = true; } else { S.rr = cc; S.lcpRP = i; } }
      else { if (cc == S.rr - 1) { bnd = S.rr; fin = true; } else { S.l = cc; S.lcpLP = i; } }
      if (!fin && ++S.guard >= 80) { fin = true; bnd = S.ubIn; }
      if (fin) {
        if (S.pass == 1) { S.b0 = bnd; S.pass = 2; S.l = S.b0 - 1; S.rr = S.ubIn; S.lcpLP = k; S.lcpRP = k; S.guard = 0; }
        else { S.b1 = bnd; if (S.b0 == S.b1) ++S.b1; S.st = LST_EXTDONE; }
      }
    }
  };

  // ---- every slot starts idle, queued as a LOOKUP item: its first step does nothing but take a read from the batch.
  // Three queue sets rotate (input of this round, output of this round, the one being cleared for the next round), so a
  // round needs a single block barrier.
  if (threadIdx.x < 9) ctl[threadIdx.x] = threadIdx.x == 0 ? static_cast<uint32_t>(SLOTS) : 0u;
  for (int s = threadIdx.x; s < SLOTS; s += NT) {
    stW[2 * SLOTS + s] = static_cast<uint32_t>(LST_IDLE) << 16;
    qBuf[s] = static_cast<uint16_t>(s);
  }
  for (int in = 0;; in = in == 2 ? 0 : in + 1) {
    __syncthreads();  // every push into queue set `in` is done, every read of the set before it too
    const int out = in == 2 ? 0 : in + 1, clr = out == 2 ? 0 : out + 1;
    const uint32_t nB = ctl[in * 2 + 0], nD = ctl[in * 2 + 1];
    if (nB + nD == 0u) break;
    if (threadIdx.x == 0) { ctl[clr * 2 + 0] = 0u; ctl[clr * 2 + 1] = 0u; ctl[6 + clr] = 0u; }
    const uint32_t chunksB = (nB + 31u) >> 5, chunksD = (nD + 31u) >> 5;
    for (;;) {
      uint32_t c = 0;
      if (lane == 0) c = atomicAdd(&ctl[6 + in], 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= chunksB + chunksD) break;
      const bool isB = c < chunksB;
      const uint32_t idx = (isB ? c : c - chunksB) * 32u + static_cast<uint32_t>(lane);
      const bool have = idx < (isB ? nB : nD);
      const int s = have ? static_cast<int>(qBuf[(in * 2 + (isB ? 0 : 1)) * SLOTS + idx]) : 0;
      RgState S;
      rgLoad<SLOTS>(stW, s, S);
      if (isB) lookupStep(S, s, have);  // every lane goes in: the step re-converges the warp before the table phase
      else if (have) probeStep(S, s);
      if (have) closure(S, s);
      __syncwarp();
      finishStep(S, s, have, out);
    }
  }
}

} // namespace rapmap_b200
