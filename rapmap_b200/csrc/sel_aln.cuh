// Kernel 4 — selective alignment: scoring of the surviving candidates and the score filter.
//
// Replaces, for a whole chunk, the `-s` block of processReadsPairSA / processReadsSingleSA (reference
// src/RapMapSAMapper.cpp:553-683, :248-330): selective_alignment::utils::getAlnScore
// (include/SelectiveAlignmentUtils.hpp:260-373: PERFECT shortcut, overhang trim, BT2 policy, UNGAPPED linear
// score, per-read alignment cache), KSW2Aligner::operator()(EXTENSION) (src/ksw2pp/KSW2Aligner.cpp:205-234) and
// ksw_extz2_sse41 in score-only mode (src/ksw2pp/ksw2_extz2_sse.c:18-304), then the minScoreFrac threshold, the
// dovetail veto, the soft / hard filter and alnScore_.
//
// Stages (all on the mapper's stream):
//   selaln_prepare_kernel  thread per pair: classifies every (hit, read end) task, emulates the alignment cache
//                          ("first earlier hit of this read end with a byte-identical reference window wins", which
//                          is what a MetroHash64-keyed map gives up to 2^-64 collisions), scores UNGAPPED tasks
//                          inline and queues the remaining ones as DP jobs.
//   ksw_extz_pair_kernel   the main DP path (ksw_pair.cuh): two jobs of one geometry per thread, int8 lanes of the SSE
//                          code as 16-bit halves of one register per band column, native 16x2 instructions; a second
//                          launch pairs the jobs that found no partner of their geometry in their tile.
//   ksw_extz_lane_kernel   thread per DP job, byte lanes of 32-bit registers, exact H[] sweep: takes what the pair kernel
//                          hands over (odd geometries; pairs whose x side left the int8 range: never seen).
//   ksw_extz_kernel        warp per DP job, any geometry: lane-for-lane restatement of the SSE kernel - int8 wrapping
//                          lanes with unsigned max/min clamps, the 16-lane block rounding that widens the band, stale
//                          out-of-band lanes, the separate int32 H[] track - in shared memory laid out exactly as
//                          the reference's kcalloc block (u|v|x|y|s|sf|qr) because its 16-byte loads and stores run
//                          past tlen/qlen into the neighbouring arrays.
//   selaln_score_kernel    thread per pair: per-hit score, best score, survivor count.
//   selaln_write_kernel    thread per pair: compacts the survivors (after an exclusive scan) with aln_score set.
#pragma once
#include <string>

#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "kernels.cuh"
#include "kmer_utils.cuh"
#include "ksw_pair.cuh"
#include "pack_swar.cuh"

namespace rapmap_b200 {

static constexpr int32_t kKswNegInf = -0x40000000;
static constexpr int32_t kIntMin = static_cast<int32_t>(0x80000000u);

struct DPJob {
  int64_t tpos;       // global text position of the target window
  uint32_t slot;      // task slot (2 * hit + end) receiving the score
  uint32_t read;      // read index in the batch view (mate 2 reads follow mate 1 reads)
  int32_t tlen1;      // target window length
  int32_t rlen;       // query length after overhang trimming
  int32_t rskip;      // bases trimmed from the query start (pos < 0)
  uint8_t rc;         // query is the reverse complement of the read
  uint8_t pad[3];
};

struct SelAlnWork {
  int32_t* taskScore{nullptr};   // [2 * hitsCap]
  int32_t* taskRef{nullptr};     // -1: own score, >= 0: copy of that slot (alignment cache hit)
  uint64_t* taskHash{nullptr};   // window hash, 0 = task never reaches the cache stage
  DPJob* jobs{nullptr};
  uint32_t* slowList{nullptr};   // DP jobs the thread-per-job kernel leaves to the general warp kernel
  uint32_t* pairLeft{nullptr};   // DP jobs without a partner of their geometry in their tile (pair kernel, second pass)
  uint32_t* exactList{nullptr};  // DP jobs the pair kernel hands to the byte-exact thread-per-job kernel
  uint32_t* jobCursor{nullptr};  // [0] job count, [1] slow-list count, [2] pairLeft count, [3] exactList count
  int32_t* hitScore{nullptr};    // [hitsCap] final per-hit score (INT_MIN = dropped)
  int32_t* pairBest{nullptr};    // [maxBatch]
  uint32_t* outCount{nullptr};   // [maxBatch + 1]
  uint64_t* outOff{nullptr};     // [maxBatch + 1]
  uint64_t hitsCap{0};
  uint64_t maxBatch{0};
  uint32_t maxReadLen{0};
};

inline void selAlnFree(SelAlnWork& w) {
  cudaFree(w.taskScore); cudaFree(w.taskRef); cudaFree(w.taskHash); cudaFree(w.jobs); cudaFree(w.slowList); cudaFree(w.pairLeft); cudaFree(w.exactList); cudaFree(w.jobCursor); cudaFree(w.hitScore);
  cudaFree(w.pairBest); cudaFree(w.outCount); cudaFree(w.outOff);
  w = SelAlnWork();
}

inline cudaError_t selAlnReserve(SelAlnWork& w, uint64_t hits) {
  if (hits <= w.hitsCap) return cudaSuccess;
  cudaFree(w.taskScore); cudaFree(w.taskRef); cudaFree(w.taskHash); cudaFree(w.jobs); cudaFree(w.slowList); cudaFree(w.pairLeft); cudaFree(w.exactList); cudaFree(w.hitScore);
  w.pairLeft = nullptr; w.exactList = nullptr; w.slowList = nullptr; w.taskScore = nullptr; w.taskRef = nullptr; w.taskHash = nullptr; w.jobs = nullptr; w.hitScore = nullptr;
  uint64_t cap = hits;
  cudaError_t e;
  if ((e = cudaMalloc(&w.taskScore, cap * 2 * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.taskRef, cap * 2 * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.taskHash, cap * 2 * 8)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.jobs, cap * 2 * sizeof(DPJob))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.slowList, cap * 2 * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.pairLeft, cap * 2 * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.exactList, cap * 2 * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.hitScore, cap * 4)) != cudaSuccess) return e;
  w.hitsCap = cap;
  return cudaSuccess;
}

inline cudaError_t selAlnAlloc(SelAlnWork& w, uint64_t maxBatch, uint32_t maxReadLen, uint64_t hitsCap) {
  w.maxBatch = maxBatch;
  w.maxReadLen = maxReadLen;
  cudaError_t e;
  if ((e = cudaMalloc(&w.jobCursor, 32)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.pairBest, maxBatch * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.outCount, (maxBatch + 1) * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&w.outOff, (maxBatch + 1) * 8)) != cudaSuccess) return e;
  return selAlnReserve(w, hitsCap);
}

struct SelAlnParams {
  DeviceIndex ix;
  DevOpts opts;
  BatchView reads;
  uint64_t numPairs;
  uint8_t pairedInput;
  const rapmap_hit_t* hits;
  const uint64_t* pairOff;
  int32_t* taskScore;
  int32_t* taskRef;
  uint64_t* taskHash;
  DPJob* jobs;
  uint32_t* jobCursor;
  int32_t* hitScore;
  int32_t* pairBest;
  uint32_t* outCount;
  const uint64_t* outOff;
  rapmap_hit_t* outHits;
  uint32_t maxReadLen;
  uint64_t hitsCap;        // records the hit / task arrays hold: a batch whose merge produced more is skipped here (the host grows and re-runs)
};

__device__ __forceinline__ void readSpan(const BatchView& b, uint64_t r, const uint8_t*& p, uint32_t& len) {
  const int mate = r >= b.n ? 1 : 0;
  const uint64_t i = r - static_cast<uint64_t>(mate) * b.n;
  if (b.off[mate]) { uint64_t o0 = b.off[mate][i]; p = b.seq[mate] + o0; len = static_cast<uint32_t>(b.off[mate][i + 1] - o0); }
  else { p = b.seq[mate] + i * b.fixedLen; len = b.fixedLen; }
}

// Query character i of the (possibly reverse-complemented) read, as the reference passes it to getAlnScore:
// raw bytes for the forward read, reverseRead() output for the reverse complement.
__device__ __forceinline__ uint8_t queryChar(const uint8_t* read, uint32_t len, bool rc, int32_t i) {
  return rc ? rcChar(__ldg(read + (len - 1 - i))) : __ldg(read + i);
}

struct TaskGeom {
  int64_t tpos;
  int32_t tlen1, rlen, rskip, keyLen;
  bool ungapped;
};

// Front half of getAlnScore (:268-312).  Returns 0 = needs a score (geometry in g), 1 = final score in `done`.
__device__ __forceinline__ int classifyTask(const DeviceIndex& ix, const DevOpts& o, uint32_t tid, int32_t pos, int32_t rlenFull, uint8_t chainStat,
                                            int32_t maxScore, TaskGeom& g, int32_t& done) {
  if (chainStat == 0) { done = maxScore; return 1; }  // PERFECT
  const int32_t tlen = __ldg(ix.txpLens + tid);
  int32_t rlen = rlenFull, rskip = 0;
  const bool invalidStart = pos < 0;
  const bool invalidEnd = (pos + rlen >= tlen);
  if (invalidStart) { rskip = -pos; rlen += pos; pos = 0; }
  if ((invalidStart || invalidEnd) && (o.alignmentPolicy == 1 || o.alignmentPolicy == 2)) { done = kIntMin; return 1; }
  if (!(pos < tlen)) { done = kIntMin; return 1; }
  const bool doUngapped = (!invalidStart) && (chainStat == 1);
  const uint32_t buf = doUngapped ? 0u : 20u;
  const uint32_t lnobuf = static_cast<uint32_t>(tlen - pos);
  const uint32_t lbuf = static_cast<uint32_t>(rlen) + buf;   // NB: unsigned, like the reference (rlen may be <= 0 after trimming)
  const bool useBuf = lbuf < lnobuf;
  const uint32_t tlen1 = lbuf < lnobuf ? lbuf : lnobuf;
  g.tpos = static_cast<int64_t>(__ldg(ix.txpOffsets + tid)) + pos;
  g.tlen1 = static_cast<int32_t>(tlen1);
  g.rlen = rlen;
  g.rskip = rskip;
  g.keyLen = static_cast<int32_t>(useBuf ? tlen1 - buf : tlen1);
  g.ungapped = doUngapped;
  return 0;
}

// 8 text bytes at any alignment (the text section is padded, reads past its end are in bounds).
__device__ __forceinline__ uint64_t textLoad8(const uint8_t* p) {
  const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(7));
  const unsigned sh = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p) & 7) * 8;
  const uint64_t lo = __ldg(a), hi = __ldg(a + 1);
  return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}

// Hash of a reference window, 8 bytes per step.  It only has to be a function of (bytes, length): equal hashes are
// confirmed byte by byte before a cached score is reused.
__device__ __forceinline__ uint64_t windowHash(const uint8_t* t, int32_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ULL ^ static_cast<uint64_t>(static_cast<uint32_t>(n));
  for (int32_t i = 0; i < n; i += 8) {
    uint64_t w = textLoad8(t + i);
    if (n - i < 8) w &= (1ULL << (8 * (n - i))) - 1ULL;
    h = (h ^ w) * 0xff51afd7ed558ccdULL;
    h ^= h >> 29;
  }
  return mix64(h) | 1ULL;  // never 0
}

__device__ __forceinline__ bool windowsEqual(const uint8_t* x, const uint8_t* y, int32_t n) {
  for (int32_t i = 0; i < n; i += 8) {
    uint64_t d = textLoad8(x + i) ^ textLoad8(y + i);
    if (n - i < 8) d &= (1ULL << (8 * (n - i))) - 1ULL;
    if (d) return false;
  }
  return true;
}

// One THREAD per pair (a pair has ~3.4 hits: with a warp per pair and a lane per hit 4 of 32 threads were active,
// profiles/r01g).  Pass 1 classifies every (hit, read end) task and hashes its reference window, pass 2 resolves the
// alignment cache against the earlier tasks of the same read end, scores UNGAPPED tasks and queues the DP jobs.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) selaln_prepare_kernel(SelAlnParams P) {
  const uint64_t gt = static_cast<uint64_t>(blockIdx.x) * (WARPS * 32) + threadIdx.x;
  const DevOpts& o = P.opts;
  const int32_t a = static_cast<int8_t>(o.ma), b = static_cast<int8_t>(o.mm);
  if (P.pairOff[P.numPairs] > P.hitsCap) return;  // merge output did not fit: nothing to score in this attempt
  for (uint64_t pi = gt; pi < P.numPairs; pi += static_cast<uint64_t>(gridDim.x) * (WARPS * 32)) {
    const uint64_t h0 = P.pairOff[pi], h1 = P.pairOff[pi + 1];
    const uint32_t cnt = static_cast<uint32_t>(h1 - h0);
    if (cnt == 0) continue;
    const bool multiMapping = cnt > 1;
    for (int end = 0; end < 2; ++end) {
      const uint64_t r = end ? P.numPairs + pi : pi;
      if (end == 1 && !P.pairedInput) break;
      const uint8_t* read;
      uint32_t rl;
      readSpan(P.reads, r, read, rl);
      const int32_t maxScore = a * static_cast<int32_t>(rl);
      // ---- pass 1: classify + hash the window
      for (uint32_t i = 0; i < cnt; ++i) {
        const rapmap_hit_t h = P.hits[h0 + i];
        const uint64_t slot = 2 * (h0 + i) + end;
        const bool paired = h.mate_status == 3;
        const bool mine = paired || (end == 0 ? (h.mate_status == 1 || h.mate_status == 0) : h.mate_status == 2);
        int32_t score = kIntMin;
        uint64_t hash = 0;
        if (mine) {
          const int32_t pos = (end == 0 || !paired) ? h.pos : h.mate_pos;
          const uint8_t cs = end == 0 ? (h.chain_status & 15) : (h.chain_status >> 4);
          TaskGeom g;
          if (classifyTask(P.ix, o, h.tid, pos, static_cast<int32_t>(rl), cs, maxScore, g, score) == 0) hash = windowHash(P.ix.text + g.tpos, g.keyLen);
        }
        P.taskScore[slot] = score;
        P.taskHash[slot] = hash;
        P.taskRef[slot] = -1;
      }
      // ---- pass 2: alignment cache = first earlier task of this read end with an identical window (:320-333,:362-368);
      //      the cache only fills for multi-mapping reads.  Then score or queue the tasks that own their result.
      for (uint32_t i = 0; i < cnt; ++i) {
        const uint64_t slot = 2 * (h0 + i) + end;
        const uint64_t hash = P.taskHash[slot];
        if (hash == 0) continue;
        const rapmap_hit_t h = P.hits[h0 + i];
        const bool paired = h.mate_status == 3;
        const int32_t pos = (end == 0 || !paired) ? h.pos : h.mate_pos;
        const bool fwd = (end == 0 || !paired) ? h.fwd : h.mate_fwd;
        const uint8_t cs = end == 0 ? (h.chain_status & 15) : (h.chain_status >> 4);
        TaskGeom g;
        int32_t dummy;
        classifyTask(P.ix, o, h.tid, pos, static_cast<int32_t>(rl), cs, maxScore, g, dummy);
        int32_t ref = -1;
        if (multiMapping) {
          for (uint32_t j = 0; j < i && ref < 0; ++j) {
            const uint64_t sj = 2 * (h0 + j) + end;
            if (P.taskHash[sj] != hash) continue;
            const rapmap_hit_t hj = P.hits[h0 + j];
            const bool pj = hj.mate_status == 3;
            const int32_t posj = (end == 0 || !pj) ? hj.pos : hj.mate_pos;
            const uint8_t csj = end == 0 ? (hj.chain_status & 15) : (hj.chain_status >> 4);
            TaskGeom gj;
            classifyTask(P.ix, o, hj.tid, posj, static_cast<int32_t>(rl), csj, maxScore, gj, dummy);
            if (gj.keyLen != g.keyLen) continue;
            if (windowsEqual(P.ix.text + g.tpos, P.ix.text + gj.tpos, g.keyLen)) ref = static_cast<int32_t>(sj);
          }
        }
        if (ref >= 0) { P.taskRef[slot] = ref; continue; }
        if (g.ungapped) {  // ungappedAln (:273-283): raw character compare, N on either side counts as a match
          const int32_t alnLen = g.rlen < g.tlen1 ? g.rlen : g.tlen1;
          int32_t sc = 0;
          for (int32_t c = 0; c < alnLen; ++c) {
            uint8_t c1 = __ldg(P.ix.text + g.tpos + c);
            const uint8_t c2 = queryChar(read, rl, !fwd, c);
            c1 = (c1 == 'N' || c2 == 'N') ? c2 : c1;
            sc += (c1 == c2) ? a : b;
          }
          P.taskScore[slot] = sc;
        } else {
          const uint32_t jx = atomicAdd(P.jobCursor, 1u);
          DPJob jb;
          jb.tpos = g.tpos; jb.slot = static_cast<uint32_t>(slot); jb.read = static_cast<uint32_t>(r); jb.tlen1 = g.tlen1; jb.rlen = g.rlen; jb.rskip = g.rskip;
          jb.rc = fwd ? 0 : 1; jb.pad[0] = jb.pad[1] = jb.pad[2] = 0;
          P.jobs[jx] = jb;
        }
      }
    }
  }
}

__device__ __forceinline__ uint8_t nt4(uint8_t c) {  // seq_nt4_table_loc, src/ksw2pp/KSW2Aligner.cpp:61-72
  if (c < 4) return c;
  switch (c | 0x20) {
    case 'a': return 0;
    case 'c': return 1;
    case 'g': return 2;
    case 't': return 3;
    default: return 4;
  }
}

struct KswParams {
  DeviceIndex ix;
  BatchView reads;
  const DPJob* jobs;
  const uint32_t* jobCount;
  int32_t* taskScore;
  uint32_t warpSmemBytes;
  int32_t tl16max;      // bytes per DP byte array (multiple of 16) for the largest window
  int8_t mat0, mat1, matN;  // match, mismatch, wildcard scores of the 5x5 matrix (KSW2Aligner.cpp:74-96)
  int8_t q, e;
  int32_t w;
  const uint32_t* jobIdx;   // nullptr: jobs[0 .. *jobCount); else the jobs listed here (the ones the lane kernel left)
  uint32_t* slowList;       // lane kernel: jobs it does not take
  uint32_t* slowCount;
  // pair kernel (two jobs of one geometry per thread): its input list (nullptr = all jobs), the jobs that found no partner
  // of their geometry in their tile (second pass), and the jobs it hands to the byte-exact thread-per-job kernel
  const uint32_t* pairSrc;
  const uint32_t* pairSrcCount;
  uint32_t* pairLeft;
  uint32_t* pairLeftCount;
  uint32_t* exactList;
  uint32_t* exactCount;
  uint32_t pairPass;
};

// General path: every geometry (any bandwidth, short windows).  State in shared memory, laid out as the reference's
// kcalloc block.
__device__ __noinline__ int32_t kswGeneral(const KswParams& P, const DPJob& jb, uint8_t* mem, int lane) {
  const int8_t q = P.q, e = P.e;
  const int qe = q + e;
  const int8_t qe2 = static_cast<int8_t>((q + e) * 2);
  const uint8_t maxSc = static_cast<uint8_t>(static_cast<int8_t>(P.mat0 + (q + e) * 2));
  {
    const int qlen = jb.rlen, tlen = jb.tlen1;
    int32_t mqe = kKswNegInf, mte = kKswNegInf;
    // early returns of ksw_extz2_sse (:60,:83): empty input, or mismatch penalty beyond 2(q+e)
    int minSc = P.mat1 < P.matN ? P.mat1 : P.matN;
    minSc = minSc < P.mat0 ? minSc : P.mat0;
    if (qlen <= 0 || tlen <= 0 || -minSc > 2 * (q + e)) {
      return kKswNegInf;
    }
    const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
    const int tl16 = tlen_ * 16;
    uint8_t* u = mem;
    uint8_t* v = u + tl16;
    uint8_t* x = v + tl16;
    uint8_t* y = x + tl16;
    uint8_t* s = y + tl16;
    uint8_t* sf = s + tl16;
    uint8_t* qr = sf + tl16;
    const int memBytes = (tlen_ * 6 + qlen_ + 1) * 16;
    int32_t* H = reinterpret_cast<int32_t*>(mem + P.warpSmemBytes - static_cast<uint32_t>(P.tl16max) * 4);
    __syncwarp();
    for (int i = lane; i < memBytes; i += 32) mem[i] = 0;   // kcalloc
    for (int i = lane; i < tl16; i += 32) H[i] = kKswNegInf;
    __syncwarp();
    const uint8_t* read;
    uint32_t rl;
    readSpan(P.reads, jb.read, read, rl);
    // query of the alignment = (rc ? reverseRead(read) : read)[rskip ...]; qr holds it reversed (:97)
    for (int t = lane; t < qlen; t += 32) qr[t] = nt4(queryChar(read, rl, jb.rc != 0, jb.rskip + (qlen - 1 - t)));
    for (int t = lane; t < tlen; t += 32) sf[t] = nt4(__ldg(P.ix.text + jb.tpos + t));
    __syncwarp();
    int w = P.w;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    int last_st = -1, last_en = -1;
    for (int r = 0; r < qlen + tlen - 1; ++r) {
      int st = 0, en = tlen - 1;
      if (st < r - qlen + 1) st = r - qlen + 1;
      if (en > r) en = r;
      if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
      if (en > ((r + w) >> 1)) en = (r + w) >> 1;
      if (st > en) break;  // zdropped (:111-114)
      const int st0 = st, en0 = en;
      st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
      int8_t x1, v1;
      if (st > 0) {
        if (st - 1 >= last_st && st - 1 <= last_en) { x1 = static_cast<int8_t>(x[st - 1]); v1 = static_cast<int8_t>(v[st - 1]); }
        else { x1 = 0; v1 = 0; }
      } else { x1 = 0; v1 = r ? q : 0; }
      __syncwarp();
      if (en >= r && lane == 0) { y[r] = 0; u[r] = static_cast<uint8_t>(r ? q : 0); }
      __syncwarp();
      // scores, whole 16-lane blocks from st0 (:126-140); two blocks per step, reads before writes
      const uint8_t* qrr = qr + (qlen - 1 - r);
      const int sEnd = st0 + ((en0 - st0) / 16 + 1) * 16;
      for (int t0 = st0; t0 < sEnd; t0 += 32) {
        const int t = t0 + lane;
        uint8_t sc = 0;
        const bool on = t < sEnd;
        if (on) {
          const uint8_t a1 = sf[t], a2 = qrr[t];
          int8_t z = (a1 == a2) ? P.mat0 : P.mat1;
          if (a1 == 4 || a2 == 4) z = P.matN;
          sc = static_cast<uint8_t>(z);
        }
        __syncwarp();
        if (on) s[t] = sc;
        __syncwarp();
      }
      // core loop (:147-164): every lane reads previous-row state, then all write
      int8_t carryX = x1, carryV = v1;
      for (int t0 = st; t0 <= en; t0 += 32) {
        const int t = t0 + lane;
        const bool on = t <= en;
        int8_t xt1 = 0, vt1 = 0, ut = 0, yt = 0, sv = 0, xo = 0, vo = 0;
        if (on) {
          xo = static_cast<int8_t>(x[t]); vo = static_cast<int8_t>(v[t]);
          ut = static_cast<int8_t>(u[t]); yt = static_cast<int8_t>(y[t]); sv = static_cast<int8_t>(s[t]);
        }
        xt1 = static_cast<int8_t>(__shfl_up_sync(0xffffffffu, static_cast<int>(xo), 1));
        vt1 = static_cast<int8_t>(__shfl_up_sync(0xffffffffu, static_cast<int>(vo), 1));
        if (lane == 0) { xt1 = carryX; vt1 = carryV; }
        carryX = static_cast<int8_t>(__shfl_sync(0xffffffffu, static_cast<int>(xo), 31));
        carryV = static_cast<int8_t>(__shfl_sync(0xffffffffu, static_cast<int>(vo), 31));
        __syncwarp();
        if (on) {
          int8_t z = static_cast<int8_t>(sv + qe2);
          int8_t aa = static_cast<int8_t>(xt1 + vt1);
          int8_t bb = static_cast<int8_t>(yt + ut);
          z = z > aa ? z : aa;                                    // _mm_max_epi8
          uint8_t zu = static_cast<uint8_t>(z), bu = static_cast<uint8_t>(bb);
          zu = zu > bu ? zu : bu;                                 // _mm_max_epu8
          zu = zu < maxSc ? zu : maxSc;                           // _mm_min_epu8
          z = static_cast<int8_t>(zu);
          u[t] = static_cast<uint8_t>(static_cast<int8_t>(z - vt1));
          v[t] = static_cast<uint8_t>(static_cast<int8_t>(z - ut));
          z = static_cast<int8_t>(z - q);
          aa = static_cast<int8_t>(aa - z);
          bb = static_cast<int8_t>(bb - z);
          x[t] = static_cast<uint8_t>(aa > 0 ? aa : 0);
          y[t] = static_cast<uint8_t>(bb > 0 ? bb : 0);
        }
        __syncwarp();
      }
      // exact max track (:228-272)
      if (r > 0) {
        const int32_t hPrev = en0 > 0 ? H[en0 - 1] : 0;
        const int32_t hSelf = H[en0];
        __syncwarp();
        for (int t0 = st0; t0 < en0; t0 += 32) {
          const int t = t0 + lane;
          if (t < en0) H[t] += static_cast<int32_t>(v[t]) - qe;
        }
        if (lane == 0) H[en0] = en0 > 0 ? hPrev + static_cast<int32_t>(u[en0]) - qe : hSelf + static_cast<int32_t>(v[en0]) - qe;
      } else {
        if (lane == 0) H[0] = static_cast<int32_t>(v[0]) - qe - qe;
      }
      __syncwarp();
      if (en0 == tlen - 1 && H[en0] > mte) mte = H[en0];
      if (r - st0 == qlen - 1 && H[st0] > mqe) mqe = H[st0];
      last_st = st; last_en = en;
    }
    __syncwarp();
    return mqe > mte ? mqe : mte;
  }
}


// Fast path for the usual geometry (0 <= w <= 15, window >= 64): at most 32 DP lanes are live on an anti-diagonal, so
// lane (t & 31) keeps u/v/x/y/s/H of column t in registers; neighbours come through shuffles, a lane whose column leaves
// the 32-wide window restarts from the zero-initialised state of its next column (t + 32).  Same cell arithmetic and the
// same stale-lane behaviour as the general path, ~4x fewer instructions and no shared-memory traffic but the codes.
__device__ __forceinline__ int32_t kswBand32(const KswParams& P, const DPJob& jb, uint8_t* mem, int lane) {
  const int qlen = jb.rlen, tlen = jb.tlen1;
  const int8_t q = P.q, e = P.e;
  const int qe = q + e;
  const int8_t qe2 = static_cast<int8_t>((q + e) * 2);
  const uint8_t maxSc = static_cast<uint8_t>(static_cast<int8_t>(P.mat0 + (q + e) * 2));
  const int tl16 = (tlen + 15) / 16 * 16, ql16 = (qlen + 15) / 16 * 16 + 16;
  uint8_t* sf = mem;
  uint8_t* qr = mem + tl16;
  __syncwarp();
  for (int i = lane; i < tl16 + ql16; i += 32) mem[i] = 0;
  __syncwarp();
  const uint8_t* read;
  uint32_t rl;
  readSpan(P.reads, jb.read, read, rl);
  for (int t = lane; t < qlen; t += 32) qr[t] = nt4(queryChar(read, rl, jb.rc != 0, jb.rskip + (qlen - 1 - t)));
  for (int t = lane; t < tlen; t += 32) sf[t] = nt4(__ldg(P.ix.text + jb.tpos + t));
  __syncwarp();
  const int w = P.w;
  int8_t u = 0, v = 0, x = 0, y = 0, s = 0;
  int32_t H = kKswNegInf, Hleft = kKswNegInf;
  int32_t mqe = kKswNegInf, mte = kKswNegInf;
  int last_st = -1, last_en = -1, curSt = 0;
  const unsigned FULL = 0xffffffffu;
  for (int r = 0; r < qlen + tlen - 1; ++r) {
    int st = 0, en = tlen - 1;
    if (st < r - qlen + 1) st = r - qlen + 1;
    if (en > r) en = r;
    if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
    if (en > ((r + w) >> 1)) en = (r + w) >> 1;
    if (st > en) break;
    const int st0 = st, en0 = en;
    st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
    // boundary inputs (previous-row state, before any lane is recycled)
    int8_t x1, v1;
    if (st > 0) {
      const int8_t xs = static_cast<int8_t>(__shfl_sync(FULL, static_cast<int>(x), (st - 1) & 31));
      const int8_t vs = static_cast<int8_t>(__shfl_sync(FULL, static_cast<int>(v), (st - 1) & 31));
      const bool have = (st - 1 >= last_st && st - 1 <= last_en);
      x1 = have ? xs : 0; v1 = have ? vs : 0;
    } else { x1 = 0; v1 = r ? q : 0; }
    if (st != curSt) {  // the window moved one 16-lane block to the right
      Hleft = __shfl_sync(FULL, H, (st - 1) & 31);
      const int tOld = curSt + ((lane - curSt) & 31);
      if (tOld < st) { u = 0; v = 0; x = 0; y = 0; s = 0; H = kKswNegInf; }
      curSt = st;
    }
    const int t = st + ((lane - st) & 31);
    if (en >= r && t == r) { y = 0; u = r ? q : 0; }
    const int sEnd = st0 + ((en0 - st0) / 16 + 1) * 16;
    if (t >= st0 && t < sEnd) {
      const uint8_t a1 = sf[t], a2 = qr[qlen - 1 - r + t];
      int8_t z = (a1 == a2) ? P.mat0 : P.mat1;
      if (a1 == 4 || a2 == 4) z = P.matN;
      s = z;
    }
    int8_t xt1 = static_cast<int8_t>(__shfl_sync(FULL, static_cast<int>(x), (lane + 31) & 31));
    int8_t vt1 = static_cast<int8_t>(__shfl_sync(FULL, static_cast<int>(v), (lane + 31) & 31));
    if (t == st) { xt1 = x1; vt1 = v1; }
    if (t <= en) {
      const int8_t ut = u;
      int8_t z = static_cast<int8_t>(s + qe2);
      int8_t aa = static_cast<int8_t>(xt1 + vt1);
      int8_t bb = static_cast<int8_t>(y + ut);
      z = z > aa ? z : aa;
      uint8_t zu = static_cast<uint8_t>(z), bu = static_cast<uint8_t>(bb);
      zu = zu > bu ? zu : bu;
      zu = zu < maxSc ? zu : maxSc;
      z = static_cast<int8_t>(zu);
      u = static_cast<int8_t>(z - vt1);
      v = static_cast<int8_t>(z - ut);
      z = static_cast<int8_t>(z - q);
      aa = static_cast<int8_t>(aa - z);
      bb = static_cast<int8_t>(bb - z);
      x = aa > 0 ? aa : 0;
      y = bb > 0 ? bb : 0;
    }
    if (r > 0) {
      int32_t hp = __shfl_sync(FULL, H, (en0 - 1) & 31);
      if (en0 - 1 < st) hp = Hleft;
      if (t >= st0 && t < en0) H += static_cast<int32_t>(static_cast<uint8_t>(v)) - qe;
      if (t == en0) H = en0 > 0 ? hp + static_cast<int32_t>(static_cast<uint8_t>(u)) - qe : H + static_cast<int32_t>(static_cast<uint8_t>(v)) - qe;
    } else if (t == 0) {
      H = static_cast<int32_t>(static_cast<uint8_t>(v)) - qe - qe;
    }
    const int32_t hEn = __shfl_sync(FULL, H, en0 & 31), hSt = __shfl_sync(FULL, H, st0 & 31);
    if (en0 == tlen - 1 && hEn > mte) mte = hEn;
    if (r - st0 == qlen - 1 && hSt > mqe) mqe = hSt;
    last_st = st; last_en = en;
  }
  return mqe > mte ? mqe : mte;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ksw_extz_kernel(KswParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t* mem = smem + static_cast<size_t>(warp) * P.warpSmemBytes;
  const uint32_t nJobs = *P.jobCount;
  for (uint32_t j = blockIdx.x * WARPS + warp; j < nJobs; j += gridDim.x * WARPS) {
    const DPJob jb = P.jobs[P.jobIdx ? P.jobIdx[j] : j];
    int minSc = P.mat1 < P.matN ? P.mat1 : P.matN;
    minSc = minSc < P.mat0 ? minSc : P.mat0;
    int32_t sc;
    const bool degenerate = jb.rlen <= 0 || jb.tlen1 <= 0 || -minSc > 2 * (P.q + P.e);
    if (!degenerate && P.w >= 0 && P.w <= 15 && jb.tlen1 >= 64) sc = kswBand32(P, jb, mem, lane);
    else sc = kswGeneral(P, jb, mem, lane);
    if (lane == 0) P.taskScore[jb.slot] = sc;
    __syncwarp();
  }
}

// Thread-per-job form of the fast path (same preconditions as kswBand32: 0 <= w <= 15, window >= 64, and the code
// strips fit the thread's shared-memory strip).  ksw_extz_kernel spends ~180 warp instructions per anti-diagonal of
// ONE job (profiles/r01g: 47 G instructions, issue-bound at 70 %); here a thread owns a job, so a warp instruction
// advances 32 jobs, and the int8 lanes of the SSE code become byte lanes of 32-bit registers: the 32 live columns of
// the band (slot c = column st + c) are 8 words each of u, v, x, y, s, updated with the wrapping byte-SIMD intrinsics
// (__vadd4 / __vsub4 / __vmaxs4 / __vmaxu4 / __vminu4 = _mm_add_epi8 / _mm_sub_epi8 / _mm_max_epi8 / _mm_max_epu8 /
// _mm_min_epu8), the left neighbours x[t-1], v[t-1] come from a byte permute of the OLD words, the int32 H[] track is one
// register per column.  When the 16-aligned window start moves, the state shifts down 16 slots and the upper 16 restart
// from the kcalloc state.  Everything is statically indexed (fully unrolled).  Stale lanes, block rounding and the H[]
// track are those of kswBand32.
__device__ __forceinline__ uint32_t rep4(int8_t v) { return static_cast<uint32_t>(static_cast<uint8_t>(v)) * 0x01010101u; }

#ifndef RAPMAP_KSW_MINB
#define RAPMAP_KSW_MINB 3  // 168 registers: the 128-register build spills ~600 bytes per thread and is 1.5x slower
#endif
template <int NT, int SEQ>
__global__ void __launch_bounds__(NT, RAPMAP_KSW_MINB) ksw_extz_lane_kernel(KswParams P) {
  extern __shared__ __align__(16) uint8_t laneSeq[];
  // strip of SEQ bytes per thread, interleaved word by word: byte b at ((b >> 2) * NT + tid) * 4 + (b & 3)
  uint8_t* myB = laneSeq + static_cast<size_t>(threadIdx.x) * 4;
  const uint32_t* myW = reinterpret_cast<const uint32_t*>(laneSeq) + threadIdx.x;
  auto byteAt = [&](int b) -> uint8_t& { return myB[(static_cast<size_t>(b >> 2) * NT) * 4 + (b & 3)]; };
  const uint32_t nJobs = *P.jobCount;
  const int8_t q = P.q, e = P.e;
  const int qe = q + e;
  const uint32_t qe2x4 = rep4(static_cast<int8_t>((q + e) * 2)), qx4 = rep4(q);
  const uint32_t maxScx4 = rep4(static_cast<int8_t>(P.mat0 + (q + e) * 2));
  const uint32_t mat0x4 = rep4(P.mat0), mat1x4 = rep4(P.mat1), matNx4 = rep4(P.matN);
  int minSc = P.mat1 < P.matN ? P.mat1 : P.matN;
  minSc = minSc < P.mat0 ? minSc : P.mat0;
  const int w = P.w;
  for (uint32_t jj = blockIdx.x * NT + threadIdx.x; jj < nJobs; jj += gridDim.x * NT) {
    const uint32_t j = P.jobIdx ? P.jobIdx[jj] : jj;
    const DPJob jb = P.jobs[j];
    const int qlen = jb.rlen, tlen = jb.tlen1;
    const int tl16 = (tlen + 15) / 16 * 16, ql16 = (qlen + 15) / 16 * 16 + 16;
    const bool degenerate = qlen <= 0 || tlen <= 0 || -minSc > 2 * (q + e);
    if (degenerate || w < 0 || w > 15 || tlen < 64 || tl16 + ql16 + 36 > SEQ) {
      P.slowList[atomicAdd(P.slowCount, 1u)] = j;
      continue;
    }
    // sf (target codes) at [0, tl16), qr (reversed query codes) at [tl16, tl16 + ql16): adjacent, as in the reference block
    for (int i = 0; i < (tl16 + ql16 + 36) / 4; ++i) const_cast<uint32_t*>(myW)[static_cast<size_t>(i) * NT] = 0u;
    const uint8_t* read;
    uint32_t rl;
    readSpan(P.reads, jb.read, read, rl);
    for (int t = 0; t < qlen; ++t) byteAt(tl16 + t) = nt4(queryChar(read, rl, jb.rc != 0, jb.rskip + (qlen - 1 - t)));
    for (int t = 0; t < tlen; ++t) byteAt(t) = nt4(__ldg(P.ix.text + jb.tpos + t));

    uint32_t U[8], V[8], X[8], Y[8], S[8];
    int32_t H[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) { U[g] = 0u; V[g] = 0u; X[g] = 0u; Y[g] = 0u; S[g] = 0u; }
#pragma unroll
    for (int c = 0; c < 32; ++c) H[c] = kKswNegInf;
    int32_t Hleft = kKswNegInf, mqe = kKswNegInf, mte = kKswNegInf;
    int curSt = 0;
    for (int r = 0; r < qlen + tlen - 1; ++r) {
      int st = 0, en = tlen - 1;
      if (st < r - qlen + 1) st = r - qlen + 1;
      if (en > r) en = r;
      if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
      if (en > ((r + w) >> 1)) en = (r + w) >> 1;
      if (st > en) break;
      const int st0 = st, en0 = en;
      st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
      uint32_t xPrevW, vPrevW;  // OLD x / v words of the group to the left (byte 3 = the column just left of this group)
      if (st != curSt) {  // the window moved one 16-lane block to the right: column st - 1 was slot 15
        xPrevW = X[3]; vPrevW = V[3];
        Hleft = H[15];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          U[g] = U[g + 4]; V[g] = V[g + 4]; X[g] = X[g + 4]; Y[g] = Y[g + 4]; S[g] = S[g + 4];
          U[g + 4] = 0u; V[g + 4] = 0u; X[g + 4] = 0u; Y[g + 4] = 0u; S[g + 4] = 0u;
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) { H[c] = H[c + 16]; H[c + 16] = kKswNegInf; }
        curSt = st;
      } else if (st > 0) { xPrevW = 0u; vPrevW = 0u; }
      else { xPrevW = 0u; vPrevW = static_cast<uint32_t>(static_cast<uint8_t>(r ? q : static_cast<int8_t>(0))) << 24; }
      const int sEnd = st0 + ((en0 - st0) / 16 + 1) * 16;
      const int cR = (en >= r) ? r - st : -1, cSt0 = st0 - st, cSEnd = sEnd - st, cEn = en - st, cEn0 = en0 - st;
      // diagonal's new cell (y = 0, u = q), :343 of the general path
      const int gR = cR >> 2;
      const uint32_t mR = 0xffu << (8 * (cR & 3)), uR = static_cast<uint32_t>(static_cast<uint8_t>(r ? q : static_cast<int8_t>(0))) << (8 * (cR & 3));
      // score window [cSt0, cSEnd): whole words between a partial first and a partial last word
      const int gA = cSt0 >> 2, gB = cSEnd >> 2;
      const uint32_t loMask = 0xffffffffu << (8 * (cSt0 & 3)), hiMask = (cSEnd & 3) ? ~(0xffffffffu << (8 * (cSEnd & 3))) : 0u;
      // reversed-query bytes for slots 4g .. 4g+3 start at strip byte qB + 4g (any alignment)
      const int qB = tl16 + (qlen - 1 - r + st);
      const uint32_t qSel = 0x3210u + 0x1111u * static_cast<uint32_t>(qB & 3);
      const uint32_t* qW = myW + static_cast<size_t>(qB >> 2) * NT;
      const uint32_t* tW = myW + static_cast<size_t>(st >> 2) * NT;
      uint32_t qLo = qW[0];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t qHi = qW[static_cast<size_t>(g + 1) * NT];
        uint32_t u4 = U[g], y4 = Y[g];
        const uint32_t xOld = X[g], vOld = V[g];
        if (g == gR) { y4 &= ~mR; u4 = (u4 & ~mR) | uR; }
        uint32_t m = 0u;
        if (g >= gA && g <= gB) {
          m = 0xffffffffu;
          if (g == gA) m &= loMask;
          if (g == gB) m &= hiMask;
        }
        if (m != 0u) {
          const uint32_t a1 = tW[static_cast<size_t>(g) * NT], a2 = __byte_perm(qLo, qHi, qSel);
          const uint32_t eq = __vcmpeq4(a1, a2);                                        // 0xff where the codes agree
          const uint32_t nn = __vcmpeq4(a1 & 0x04040404u, 0x04040404u) | __vcmpeq4(a2 & 0x04040404u, 0x04040404u);  // code 4: wildcard
          uint32_t s4 = (mat0x4 & eq) | (mat1x4 & ~eq);
          s4 = (s4 & ~nn) | (matNx4 & nn);
          S[g] = (S[g] & ~m) | (s4 & m);
        }
        qLo = qHi;
        if (4 * g <= cEn) {
          const uint32_t xt1 = __byte_perm(xPrevW, xOld, 0x6543u), vt1 = __byte_perm(vPrevW, vOld, 0x6543u);
          uint32_t z = __vadd4(S[g], qe2x4);
          uint32_t aa = __vadd4(xt1, vt1);
          uint32_t bb = __vadd4(y4, u4);
          z = __vmaxs4(z, aa);      // _mm_max_epi8
          z = __vmaxu4(z, bb);      // _mm_max_epu8
          z = __vminu4(z, maxScx4); // _mm_min_epu8
          const uint32_t un = __vsub4(z, vt1), vn = __vsub4(z, u4);
          z = __vsub4(z, qx4);
          aa = __vsub4(aa, z);
          bb = __vsub4(bb, z);
          X[g] = __vmaxs4(aa, 0u);
          y4 = __vmaxs4(bb, 0u);
          u4 = un;
          V[g] = vn;
        }
        U[g] = u4; Y[g] = y4;
        xPrevW = xOld; vPrevW = vOld;
      }
      // exact max track (general path :397-412).  The sweep carries the previous column's OLD H, so H[en0] = H[en0 - 1]
      // (before this diagonal) + u - qe needs no dynamically indexed register read.
      if (r > 0) {
        int32_t hPrevOld = Hleft;  // column st - 1
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int32_t hOld = H[c];
          const int32_t vb = static_cast<int32_t>((V[c >> 2] >> (8 * (c & 3))) & 0xffu);
          int32_t hNew = hOld;
          if (c >= cSt0 && c < cEn0) hNew = hOld + vb - qe;
          if (c == cEn0) {
            const int32_t ub = static_cast<int32_t>((U[c >> 2] >> (8 * (c & 3))) & 0xffu);
            hNew = en0 > 0 ? hPrevOld + ub - qe : hOld + vb - qe;
          }
          H[c] = hNew;
          hPrevOld = hOld;
        }
      } else {
        H[0] = static_cast<int32_t>(V[0] & 0xffu) - qe - qe;
      }
      // the two read-outs happen on the last ~w + (tlen - qlen) diagonals only
      if (en0 == tlen - 1) {
        int32_t hEn = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c == cEn0) hEn = H[c];
        if (hEn > mte) mte = hEn;
      }
      if (r - st0 == qlen - 1) {
        int32_t hSt = 0;
#pragma unroll
        for (int c = 0; c < 16; ++c) if (c == cSt0) hSt = H[c];  // st0 - st < 16
        if (hSt > mqe) mqe = hSt;
      }
    }
    P.taskScore[jb.slot] = mqe > mte ? mqe : mte;
  }
}

// nt4 codes (seq_nt4_table_loc) of `len` bytes at p, eight bytes per aligned 64-bit load (only words that hold a byte of
// the range are touched), branch-free per byte; comp: the codes of rapmap::utils::reverseRead's output for these bytes
// (A<->T, C<->G, U -> A, anything else N).  put(i, code) receives byte i of the range.
template <typename Put>
__device__ __forceinline__ void fillCodes8(const uint8_t* p, int len, bool comp, Put put) {
  const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(7));
  const int off = static_cast<int>(reinterpret_cast<uintptr_t>(p) & 7);
  for (int i0 = -off; i0 < len; i0 += 32, a += 4) {
    uint64_t wd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wd[j] = (i0 + 8 * j < len) ? __ldg(a + j) : 0ULL;   // four loads in flight
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int i = i0 + 8 * j + b;
        if (i < 0 || i >= len) continue;
        put(i, baseCode(static_cast<uint32_t>(wd[j] >> (8 * b)) & 0xffu, comp));
      }
    }
  }
}

// Pair kernel: TWO DP jobs of one geometry (qlen, tlen) per thread, 16-bit halves of one register per band column
// (ksw_pair.cuh has the DP and the argument why it is exact).  A tile of 2 * NT jobs is sorted by geometry in shared
// memory (cub::BlockRadixSort, blocked: thread i gets sorted items 2i and 2i + 1); equal neighbours run as a pair, an
// item without a partner runs alone (both halves hold the same job) and its unequal neighbour goes to the `pairLeft`
// list, which a second launch (pairPass = 1) pairs up across tiles (what is still single there goes to the byte-exact kernel).  Jobs outside the pair kernel's geometry, and pairs
// whose x-side left the int8 range, go to the byte-exact thread-per-job kernel through `exactList`.
template <int NT, int CELLS>
__global__ void __launch_bounds__(NT, 2) ksw_extz_pair_kernel(KswParams P) {
  extern __shared__ __align__(16) uint8_t pairStrip[];
  using Sort = cub::BlockRadixSort<uint32_t, NT, 2, uint32_t>;
  __shared__ typename Sort::TempStorage sortTmp;
  uint32_t* myW = reinterpret_cast<uint32_t*>(pairStrip) + threadIdx.x;
  uint8_t* myB = pairStrip + static_cast<size_t>(threadIdx.x) * 4;
  // cell p (two bytes: job 0, job 1) of this thread's strip
  auto cellByte = [&](int p, int half) -> uint8_t& { return myB[(static_cast<size_t>(p >> 1) * NT) * 4 + (p & 1) * 2 + half]; };
  const kswpair::Consts C = kswpair::makeConsts(P.mat0, P.mat1, P.matN, P.q, P.e, P.w);
  const uint32_t nSrc = *P.pairSrcCount;
  constexpr uint32_t kNoJob = 0xFFFFFFu;
  for (uint32_t base = blockIdx.x * 2u * NT; base < nSrc; base += gridDim.x * 2u * NT) {
    uint32_t keys[2], vals[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t i = base + 2u * threadIdx.x + k;
      keys[k] = kNoJob; vals[k] = 0u;
      if (i < nSrc) {
        const uint32_t j = P.pairSrc ? P.pairSrc[i] : i;
        const int qlen = P.jobs[j].rlen, tlen = P.jobs[j].tlen1;
        if (kswpair::geomOk(C, qlen, tlen, CELLS)) { keys[k] = (static_cast<uint32_t>(qlen) << 11) | static_cast<uint32_t>(tlen); vals[k] = j; }
        else P.exactList[atomicAdd(P.exactCount, 1u)] = j;
      }
    }
    __syncthreads();
    Sort(sortTmp).Sort(keys, vals, 0, 24);
    int nrep = 0;
    if (keys[0] != kNoJob) {
      nrep = 1;
      if (keys[1] != keys[0] && keys[1] != kNoJob) {
        // no partner in this tile: to the second pass; still none there (a handful of jobs per batch): to the thread-per-job
        // kernel, which runs behind this one anyway (a second DP in this thread would double the tail of the launch)
        if (P.pairPass == 0) P.pairLeft[atomicAdd(P.pairLeftCount, 1u)] = vals[1];
        else P.exactList[atomicAdd(P.exactCount, 1u)] = vals[1];
      }
    }
    for (int rep = 0; rep < nrep; ++rep) {
      const bool both = keys[1] == keys[0];
      const uint32_t j0 = vals[0], j1 = both ? vals[1] : j0;
      const DPJob ja = P.jobs[j0], jb = P.jobs[j1];
      const int qlen = ja.rlen, tlen = ja.tlen1;
      const int TL = kswpair::stripTL(tlen), words = (kswpair::stripCells(qlen, tlen) + 1) / 2;
      for (int i = 0; i < words; ++i) myW[static_cast<size_t>(i) * NT] = 0u;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const DPJob& jx = half ? jb : ja;
        const uint8_t* read;
        uint32_t rl;
        readSpan(P.reads, jx.read, read, rl);
        // cell TL + i = code of query[qlen - 1 - i]: forward reads run backwards through the read, reverse-complemented ones
        // forwards through it with the complement table (reverseRead)
        const bool rc = jx.rc != 0;
        const uint8_t* qsrc = rc ? read + (static_cast<int>(rl) - jx.rskip - qlen) : read + jx.rskip;
        fillCodes8(qsrc, qlen, rc, [&](int i, uint32_t code) { cellByte(TL + (rc ? i : qlen - 1 - i), half) = static_cast<uint8_t>(code); });
        fillCodes8(P.ix.text + jx.tpos, tlen, false, [&](int t, uint32_t code) { cellByte(t, half) = static_cast<uint8_t>(code); });
      }
      int32_t sc0, sc1;
      if (kswpair::pairDP<NT>(myW, qlen, tlen, C, sc0, sc1)) {
        P.taskScore[ja.slot] = sc0;
        if (both) P.taskScore[jb.slot] = sc1;
      } else {
        P.exactList[atomicAdd(P.exactCount, 1u)] = j0;
        if (both) P.exactList[atomicAdd(P.exactCount, 1u)] = j1;
      }
    }
  }
}

__device__ __forceinline__ int32_t taskValue(const SelAlnParams& P, uint64_t slot) {
  const int32_t ref = P.taskRef[slot];
  return ref >= 0 ? P.taskScore[ref] : P.taskScore[slot];
}

// Per-hit score, best score and survivor count of one pair (src/RapMapSAMapper.cpp:567-683).
__global__ void __launch_bounds__(256) selaln_score_kernel(SelAlnParams P) {
  const DevOpts& o = P.opts;
  const int32_t a = static_cast<int8_t>(o.ma);
  const bool skip = P.pairOff[P.numPairs] > P.hitsCap;
  for (uint64_t pi = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; pi < P.numPairs; pi += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (skip) { P.outCount[pi] = 0; continue; }
    const uint64_t h0 = P.pairOff[pi], h1 = P.pairOff[pi + 1];
    if (h0 == h1) { P.outCount[pi] = 0; continue; }
    const uint8_t* rd;
    uint32_t l1 = 0, l2 = 0;
    readSpan(P.reads, pi, rd, l1);
    if (P.pairedInput) readSpan(P.reads, P.numPairs + pi, rd, l2);
    const int32_t maxLeft = a * static_cast<int32_t>(l1), maxRight = a * static_cast<int32_t>(l2);
    const double optFrac = o.minScoreFraction;
    int32_t best = kIntMin;
    for (uint64_t h = h0; h < h1; ++h) {
      const rapmap_hit_t hit = P.hits[h];
      int32_t score = kIntMin;
      if (hit.mate_status == 3) {
        int32_t s1 = taskValue(P, 2 * h), s2 = taskValue(P, 2 * h + 1);
        if (hit.fwd != hit.mate_fwd && o.noDovetail) {
          if (hit.fwd && (hit.pos > hit.mate_pos)) { s1 = kIntMin; s2 = kIntMin; }
          else if (hit.mate_fwd && (hit.mate_pos > hit.pos)) { s1 = kIntMin; s2 = kIntMin; }
        }
        if ((static_cast<double>(s1) < __dmul_rn(optFrac, static_cast<double>(maxLeft))) || (static_cast<double>(s2) < __dmul_rn(optFrac, static_cast<double>(maxRight)))) score = kIntMin;
        else score = s1 + s2;
      } else if (hit.mate_status == 2) {
        const int32_t s = taskValue(P, 2 * h + 1);
        score = (static_cast<double>(s) < __dmul_rn(optFrac, static_cast<double>(maxRight))) ? kIntMin : s;
      } else {  // PAIRED_END_LEFT or SINGLE_END
        const int32_t s = taskValue(P, 2 * h);
        score = (static_cast<double>(s) < __dmul_rn(optFrac, static_cast<double>(maxLeft))) ? kIntMin : s;
      }
      P.hitScore[h] = score;
      best = score > best ? score : best;
    }
    uint32_t keep = 0;
    if (best > kIntMin) {
      for (uint64_t h = h0; h < h1; ++h) {
        const int32_t sc = P.hitScore[h];
        const bool rem = o.hardFilter ? (sc < best) : (sc == kIntMin);
        keep += rem ? 0u : 1u;
      }
    }
    P.pairBest[pi] = best;
    P.outCount[pi] = keep;
  }
}

__global__ void __launch_bounds__(256) selaln_write_kernel(SelAlnParams P) {
  const DevOpts& o = P.opts;
  if (P.pairOff[P.numPairs] > P.hitsCap) return;
  for (uint64_t pi = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; pi < P.numPairs; pi += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t h0 = P.pairOff[pi], h1 = P.pairOff[pi + 1];
    const int32_t best = P.pairBest[pi];
    if (h0 == h1 || !(best > kIntMin)) continue;
    uint64_t w = P.outOff[pi];
    for (uint64_t h = h0; h < h1; ++h) {
      const int32_t sc = P.hitScore[h];
      const bool rem = o.hardFilter ? (sc < best) : (sc == kIntMin);
      if (rem) continue;
      rapmap_hit_t hit = P.hits[h];
      hit.aln_score = sc;
      P.outHits[w++] = hit;
    }
  }
}

struct CastU64b {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return v; }
};

// Launch geometry of the DP kernels, fixed per mapper (computed once by selAlnSetup).
struct SelAlnLaunch {
  uint32_t warpSmemBytes{0};
  int32_t tl16max{0};
  uint32_t smemK{0}, smemL{0}, smemP{0};
  int occK{1}, occL{1}, occP{1};
  bool laneKsw{true}, pairKsw{true};
};
static constexpr int kKswWarps = 4, kKswLaneThreads = 128, kKswLaneSeq = 352;
static constexpr int kKswPairThreads = 128, kKswPairCells = 384;   // strip cells per thread: reads up to ~155 bases

inline int selAlnSetup(const SelAlnWork& w, SelAlnLaunch& L, std::string& err) {
  auto cuFail = [&](const char* what, cudaError_t e) { err = std::string(what) + ": " + cudaGetErrorString(e); return RAPMAP_ERR_CUDA; };
  // DP kernel: shared memory per warp = the reference's kcalloc block for the largest window + the int32 H track
  const int tlenMax = static_cast<int>(w.maxReadLen) + 20;
  const int tl16 = (tlenMax + 15) / 16 * 16;
  const int qlen_ = (static_cast<int>(w.maxReadLen) + 15) / 16;
  L.tl16max = tl16;
  L.warpSmemBytes = static_cast<uint32_t>((tl16 / 16 * 6 + qlen_ + 1) * 16 + tl16 * 4);
  L.smemK = L.warpSmemBytes * kKswWarps;
  if (L.smemK > 200 * 1024) { err = "max_read_len too large for the ksw2 shared-memory layout"; return RAPMAP_ERR_ARG; }
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(ksw_extz_kernel<kKswWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.smemK))) != cudaSuccess)
    return cuFail("cudaFuncSetAttribute(ksw)", e);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.occK, ksw_extz_kernel<kKswWarps>, kKswWarps * 32, L.smemK);
  if (L.occK < 1) L.occK = 1;
  const char* sel = std::getenv("RAPMAP_B200_KSW");
  L.laneKsw = !(sel && std::string(sel) == "warp");
  if (L.laneKsw) {
    L.smemL = kKswLaneThreads * kKswLaneSeq;
    if ((e = cudaFuncSetAttribute(ksw_extz_lane_kernel<kKswLaneThreads, kKswLaneSeq>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.smemL))) != cudaSuccess)
      return cuFail("cudaFuncSetAttribute(ksw lane)", e);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.occL, ksw_extz_lane_kernel<kKswLaneThreads, kKswLaneSeq>, kKswLaneThreads, L.smemL);
    if (L.occL < 1) L.occL = 1;
  }
  L.pairKsw = L.laneKsw && !(sel && std::string(sel) == "lane");
  if (L.pairKsw) {
    L.smemP = kKswPairThreads * kKswPairCells * 2;
    if ((e = cudaFuncSetAttribute(ksw_extz_pair_kernel<kKswPairThreads, kKswPairCells>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.smemP))) != cudaSuccess)
      return cuFail("cudaFuncSetAttribute(ksw pair)", e);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.occP, ksw_extz_pair_kernel<kKswPairThreads, kKswPairCells>, kKswPairThreads, L.smemP);
    if (L.occP < 1) L.occP = 1;
  }
  return RAPMAP_OK;
}

// Enqueues the four stages on `st` (no host synchronisation): afterwards dPairOff holds the post-filter offsets,
// dOutHits the surviving hits, *hTotal (pinned) their number and hJobs[0..1] the DP job counts (all / left to the
// general kernel).  The work arrays must hold hitsCap records (selAlnReserve); a batch whose merge produced more is
// skipped by the kernels and re-run by the caller after growing.
inline int selAlnEnqueue(SelAlnWork& w, const SelAlnLaunch& L, const DeviceIndex& ix, const DevOpts& opts, const BatchView& bv, uint64_t n, bool paired,
                         rapmap_hit_t* dHits, rapmap_hit_t* dOutHits, uint64_t* dPairOff, uint64_t hitsCap, void* dCubTemp, size_t cubTempBytes, int numSMs, cudaStream_t st,
                         uint32_t* launches, uint64_t* hTotal, uint32_t* hJobs, cudaEvent_t evKsw0, cudaEvent_t evKsw1, std::string& err) {
  auto cuFail = [&](const char* what, cudaError_t e) { err = std::string(what) + ": " + cudaGetErrorString(e); return RAPMAP_ERR_CUDA; };
  cudaError_t e;
  SelAlnParams sp{};
  sp.ix = ix; sp.opts = opts; sp.reads = bv; sp.numPairs = n; sp.pairedInput = paired ? 1 : 0; sp.hits = dHits; sp.pairOff = dPairOff;
  sp.taskScore = w.taskScore; sp.taskRef = w.taskRef; sp.taskHash = w.taskHash; sp.jobs = w.jobs; sp.jobCursor = w.jobCursor; sp.hitScore = w.hitScore;
  sp.pairBest = w.pairBest; sp.outCount = w.outCount; sp.outOff = w.outOff; sp.outHits = dOutHits; sp.maxReadLen = w.maxReadLen;
  sp.hitsCap = hitsCap < w.hitsCap ? hitsCap : w.hitsCap;
  if ((e = cudaMemsetAsync(w.jobCursor, 0, 32, st)) != cudaSuccess) return cuFail("memset", e);
  constexpr int W = 8;
  int g = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(numSMs) * 8, (n + W * 32 - 1) / (W * 32)));
  selaln_prepare_kernel<W><<<g, W * 32, 0, st>>>(sp);
  ++*launches;
  KswParams kp{};
  kp.ix = ix; kp.reads = bv; kp.jobs = w.jobs; kp.jobCount = w.jobCursor; kp.taskScore = w.taskScore;
  kp.tl16max = L.tl16max;
  kp.warpSmemBytes = L.warpSmemBytes;
  {
    int a = opts.ma, b = opts.mm;  // KSW2Aligner ctor (:74-96)
    a = a < 0 ? -a : a;
    b = b > 0 ? -b : b;
    kp.mat0 = static_cast<int8_t>(a); kp.mat1 = static_cast<int8_t>(b); kp.matN = 0;
  }
  kp.q = static_cast<int8_t>(opts.go); kp.e = static_cast<int8_t>(opts.ge); kp.w = opts.dpBandwidth;
  if (evKsw0 && (e = cudaEventRecord(evKsw0, st)) != cudaSuccess) return cuFail("event", e);
  if (L.pairKsw) {  // two jobs of one geometry per thread; second pass over the jobs that found no partner in their tile
    kp.pairSrc = nullptr; kp.pairSrcCount = w.jobCursor; kp.pairLeft = w.pairLeft; kp.pairLeftCount = w.jobCursor + 2;
    kp.exactList = w.exactList; kp.exactCount = w.jobCursor + 3; kp.pairPass = 0;
    ksw_extz_pair_kernel<kKswPairThreads, kKswPairCells><<<numSMs * L.occP, kKswPairThreads, L.smemP, st>>>(kp);
    kp.pairSrc = w.pairLeft; kp.pairSrcCount = w.jobCursor + 2; kp.pairPass = 1;
    ksw_extz_pair_kernel<kKswPairThreads, kKswPairCells><<<numSMs * L.occP, kKswPairThreads, L.smemP, st>>>(kp);
    *launches += 2;
    kp.jobIdx = w.exactList; kp.jobCount = w.jobCursor + 3;   // the byte-exact kernel below takes what they handed over
  }
  if (L.laneKsw) {  // thread-per-job kernel; what it leaves (short windows, wide bands, long reads) goes to the warp kernel
    kp.slowList = w.slowList; kp.slowCount = w.jobCursor + 1;
    ksw_extz_lane_kernel<kKswLaneThreads, kKswLaneSeq><<<numSMs * L.occL, kKswLaneThreads, L.smemL, st>>>(kp);
    ++*launches;
    kp.jobIdx = w.slowList; kp.jobCount = w.jobCursor + 1;
  }
  ksw_extz_kernel<kKswWarps><<<numSMs * L.occK, kKswWarps * 32, L.smemK, st>>>(kp);
  ++*launches;
  if (evKsw1 && (e = cudaEventRecord(evKsw1, st)) != cudaSuccess) return cuFail("event", e);
  int g3 = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(numSMs) * 8, (n + 255) / 256));
  if ((e = cudaMemsetAsync(w.outCount + n, 0, 4, st)) != cudaSuccess) return cuFail("memset", e);
  selaln_score_kernel<<<g3, 256, 0, st>>>(sp);
  ++*launches;
  {
    cub::TransformInputIterator<uint64_t, CastU64b, uint32_t*> it(w.outCount, CastU64b());
    size_t tb = cubTempBytes;
    if ((e = cub::DeviceScan::ExclusiveSum(dCubTemp, tb, it, w.outOff, static_cast<int>(n + 1), st)) != cudaSuccess) return cuFail("scan", e);
  }
  selaln_write_kernel<<<g3, 256, 0, st>>>(sp);
  ++*launches;
  if ((e = cudaMemcpyAsync(hTotal, w.outOff + n, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuFail("memcpy", e);
  if ((e = cudaMemcpyAsync(hJobs, w.jobCursor, 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuFail("memcpy", e);
  if ((e = cudaMemcpyAsync(dPairOff, w.outOff, (n + 1) * 8, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return cuFail("memcpy", e);
  return RAPMAP_OK;
}

} // namespace rapmap_b200
