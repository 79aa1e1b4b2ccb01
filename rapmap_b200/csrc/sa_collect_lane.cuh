// Kernel 1, lane-per-read form — "the SA-lookup kernel": seed + maximal-mappable-prefix collection, ONE THREAD PER READ.
//
// Replaces SACollector::operator() / getSAHits_ / spotCheck_ (reference include/SACollector.hpp:108-362, :441-677,
// :366-431), SASearcher::extendSearchNaive and ::lce (include/SASearcher.hpp:87-334), Kmer encode / RC / homopolymer
// (include/Kmer.hpp:92-100,484-542) and the khash / FrugalBooMap find (include/RapMapUtils.hpp:65-67).
//
// Why a thread per read.  The walk of one read is a chain of dependent random reads (hash probe -> SA probe -> text
// compare) whose decision sequence has to be replayed exactly; a warp that owns one read spends 31/32 of its issue
// slots replicating scalar state and has one read's worth of memory parallelism (profiles/r01c: 5.9k warp instructions
// per read, 26 % issue utilisation, DRAM 4 % of peak).  Here every lane walks its own read, so a resident warp has 32
// independent chains in flight and an SM ~1000; HBM latency is covered by reads, not by speculation.
//
// Divergence is handled by writing the walk as a state machine with a fixed phase order per loop trip:
//   refill -> (A) advance to the next k-mer that needs a lookup -> (B) ONE pair of hash lookups (k-mer and its reverse
//   complement, loads issued together) -> (C) set up a suffix search -> (D) ONE binary-search probe (SA load + text
//   compare, 8 bytes per step) -> (E) bookkeeping / interval record / strand sequencing -> publish.
// Lanes meet again at every phase, so the two expensive phases are executed once per trip for all lanes that need them.
// A lane that finishes its read fetches a new one (warp-aggregated atomic on a read cursor) as soon as enough lanes are
// idle, so a warp's 32 chains stay occupied until the batch runs dry.
//
// The read lives in shared memory as 16-byte words of 32 bases {2-bit codes (u64), non-ACGT mask, N mask}, written by
// pack_reads_kernel (coalesced, one warp per read) and interleaved across the block (word j of thread t at
// [j * NT + t]) so a warp's accesses are conflict-free.  A k-mer at any position of either strand is two LDS.128 and a
// funnel shift; 8 query characters for the text compare are rebuilt from the codes with two PRMTs.  Windows that
// contain a non-ACGT base take exact slow paths (partial k-mer words of Kmer::fromChars, original bytes re-read).
#pragma once
#include <climits>
#include "kmer_utils.cuh"
#include "pack_swar.cuh"

namespace rapmap_b200 {

struct LaneParams {
  DeviceIndex ix;
  BatchView reads;
  DevOpts opts;
  uint32_t maxReadLen;
  uint32_t nw;             // 32-base words per read (ceil(maxReadLen / 32))
  uint4* packed;           // [numReads][nw] {codes_lo, codes_hi, invalid mask, N mask}
  uint4* kmask;            // [numReads][nw] per 32 k-mer start positions {absent fwd, absent rc, window valid, homopolymer} (kmer_mask_kernel)
  ReadSummary* summ;
  IntervalRec* arena;
  uint32_t arenaCap;
  uint32_t* arenaCursor;
  uint32_t* status;
  IntervalRec* ivScratch;  // per resident thread: forward list [0, ivStride), reverse-complement list [ivStride, 2 ivStride)
  uint32_t ivStride;
  uint32_t* voteScratch;   // per resident thread: 3 x voteWords (tested, fwd present, rc present); k-mer vote mode only
  uint32_t voteWords;
  uint32_t* readCursor;
  uint32_t maskChunks;     // 8-window chunks per read the mask kernel fills: ceil((maxReadLen - k + 1) / 8)
  uint32_t* classCtl;      // [0, 8) class histogram, [8, 16) scatter cursors (zeroed per batch)
  uint32_t* order;         // [numReads] read indices grouped by work class; the walk kernel takes reads in this order
  uint32_t masksInGlobal;  // 1: reads so long that packed words + masks exceed the shared memory of a block: the masks are read from kmask
};

// ---- K0: pack reads.  One thread per 32-base word of a read (a warp covers 8 reads x 4 words = 800 contiguous bytes).
// Codes: A C G T = 0..3 (either case); 'U' = 3 with the invalid bit set (reverseRead turns U into A, so on the
// reverse-complement strand it is a valid base, src/RapMapUtils.cpp:63-72); N = 1 + invalid; anything else 0 + invalid.
__global__ void __launch_bounds__(256) pack_reads_kernel(LaneParams P) {
  const uint64_t nWords = P.reads.numReads * P.nw;
  for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < nWords; g += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t r = g / P.nw;
    const uint32_t j = static_cast<uint32_t>(g - r * P.nw);
    const int mate = r >= P.reads.n ? 1 : 0;
    const uint64_t ri = r - static_cast<uint64_t>(mate) * P.reads.n;
    const uint8_t* src;
    uint32_t len;
    if (P.reads.off[mate]) {
      const uint64_t o0 = P.reads.off[mate][ri], o1 = P.reads.off[mate][ri + 1];
      src = P.reads.seq[mate] + o0;
      len = static_cast<uint32_t>(o1 - o0);
    } else {
      src = P.reads.seq[mate] + ri * P.reads.fixedLen;
      len = P.reads.fixedLen;
    }
    if (len > P.maxReadLen) len = 0;  // reported by the SA-lookup kernel
    const int L = static_cast<int>(len);
    const int i0 = static_cast<int>(j) * 32;
    uint64_t codes = 0;
    uint32_t inv = 0, nn = 0;
    const int n = L - i0 < 32 ? L - i0 : 32;
    const uint8_t* p = src + i0;
    if (n == 32 && (reinterpret_cast<uintptr_t>(p) & 3) == 0) {
      // full word, 4-byte aligned: four bases per 32-bit operation (pack_swar.cuh; the byte loop below costs ~33 instructions
      // per base and made this kernel math-pipe bound, profiles/r02l)
      const uint32_t* p4 = reinterpret_cast<const uint32_t*>(p);
      uint32_t x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __ldg(p4 + j);
      packBases32(x, codes, inv, nn);
    } else {
    for (int b = 0; b < n; ++b) {
      const uint32_t ch = __ldg(src + i0 + b);
      const uint32_t uc = ch & 0xDFu;
      const bool ok = uc == 'A' || uc == 'C' || uc == 'G' || uc == 'T';
      const uint32_t code = ok ? (((ch >> 1) ^ (ch >> 2)) & 3u) : (uc == 'U' ? 3u : (uc == 'N' ? 1u : 0u));
      codes |= static_cast<uint64_t>(code) << (62 - 2 * b);
      if (!ok) inv |= 1u << b;
      if (uc == 'N') nn |= 1u << b;
    }
    }
    P.packed[g] = make_uint4(static_cast<uint32_t>(codes), static_cast<uint32_t>(codes >> 32), inv, nn);
  }
}

// ---- K0b: k-mer masks.  The walk of a noisy read spends most of its steps on k-mers that are not in the index (the ~31
// windows covering a sequencing error, both orientations): each such step is "advance one base" with no other effect
// (include/SACollector.hpp:520-546,:671-673).  Whether a window is such a dead position is a pure function of the window,
// so it is computed here for EVERY window of every read, fully convergent (a thread owns 8 consecutive windows; the four
// threads of a 32-window word combine their bytes with shuffles), and the walk kernel skips runs of dead positions with a
// bit scan instead of a ~200-instruction loop trip with two dependent L2 loads each.
// Per forward window q (bit q & 31 of word q >> 5): V = the window holds only ACGT; for V windows H = homopolymer,
// AF / AR = the k-mer filter proves the k-mer / its reverse complement absent.  Windows with a non-ACGT base (V = 0) keep
// taking the walk kernel's exact path (partial words of Kmer::fromChars, the N skips, 'U' on the reverse-complement strand).
__global__ void __launch_bounds__(256) kmer_mask_kernel(LaneParams P) {
  // task = (read, chunk of 8 windows); chunks beyond the longest possible read are never touched (the mask array is zeroed
  // once when the mapper is created) and never read (a walk only looks at windows q <= L - k)
  const uint32_t cpr = P.maskChunks;  // ceil((maxReadLen - k + 1) / 8)
  const uint64_t nTasks = P.reads.numReads * cpr;
  const int k = static_cast<int>(P.ix.k);
  const uint32_t kbits = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
  const uint64_t kmask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
  uint8_t* out = reinterpret_cast<uint8_t*>(P.kmask);
  for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < nTasks; g += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t r = g / cpr;
    const int c = static_cast<int>(g - r * cpr);
    const int wi = c >> 2, sh = (c & 3) * 8;
    const uint64_t wordIdx = r * P.nw + wi;
    const int mate = r >= P.reads.n ? 1 : 0;
    const uint64_t ri = r - static_cast<uint64_t>(mate) * P.reads.n;
    uint32_t len = P.reads.fixedLen;
    if (P.reads.off[mate]) len = static_cast<uint32_t>(P.reads.off[mate][ri + 1] - P.reads.off[mate][ri]);
    if (len > P.maxReadLen) len = 0;
    const int L = static_cast<int>(len);
    const int q0 = c * 8;
    uint32_t af = 0, ar = 0, vv = 0, hh = 0;
    if (q0 + k <= L) {
      const uint4 a = __ldg(P.packed + wordIdx);
      uint4 b = make_uint4(0u, 0u, 0u, 0u);
      if (wi + 1 < static_cast<int>(P.nw)) b = __ldg(P.packed + wordIdx + 1);
      const uint64_t ca = (static_cast<uint64_t>(a.y) << 32) | a.x, cb = (static_cast<uint64_t>(b.y) << 32) | b.x;
      const uint64_t hi = sh ? ((ca << (2 * sh)) | (cb >> (64 - 2 * sh))) : ca;  // bases q0 .. q0+31
      uint64_t lo = cb << (2 * sh);                                              // bases q0+32 .. (k + 7 <= 38 bases are needed)
      const uint64_t inv = ((static_cast<uint64_t>(b.z) << 32) | a.z) >> sh;     // bit i: base q0+i is not ACGT
      // rolling k-mer and reverse complement (Kmer::shiftFw / getRC, include/Kmer.hpp:92-100); `same` counts the equal
      // neighbours inside the window: k - 1 of them = homopolymer (:484-487)
      uint64_t w = hi >> (64 - 2 * k);
      uint64_t wr = kmerRC(w, k);
      int same = 0;
      for (int j = 1; j < k; ++j) same += (((w >> (2 * j)) ^ (w >> (2 * j - 2))) & 3ULL) == 0ULL;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (q0 + i + k > L) break;
        if (i > 0) {  // window q0+i: drop base q0+i-1, take base q0+i+k-1 (bit 63:62 of the remaining stream)
          const int kk = i + k - 1;  // base index relative to q0
          const uint64_t nb = kk < 32 ? ((hi >> (62 - 2 * kk)) & 3ULL) : ((lo >> 62) & 3ULL);
          if (kk >= 32) lo <<= 2;
          same -= (((w >> (2 * k - 2)) ^ (w >> (2 * k - 4))) & 3ULL) == 0ULL;  // pair (first, second) leaves
          same += ((w ^ nb) & 3ULL) == 0ULL;                                     // pair (last, new) enters
          w = ((w << 2) | nb) & kmask;
          wr = (wr >> 2) | ((3ULL - nb) << (2 * k - 2));
        }
        if ((static_cast<uint32_t>(inv >> i) & kbits) != 0u) continue;
        const uint32_t bit = 1u << i;
        vv |= bit;
        if (same == k - 1) hh |= bit;
        if (P.ix.filter != nullptr) {
          uint64_t fw;
          uint32_t ma, mb;
          filterQuery(w, wr, P.ix.filterShift, fw, ma, mb);  // ONE probe answers both orientations (canonical key)
          const uint32_t fv = ldgKeep(P.ix.filter + fw);
          if ((fv & ma) != ma) af |= bit;
          if ((fv & mb) != mb) ar |= bit;
        }
      }
    }
    uint8_t* o = out + wordIdx * 16 + (c & 3);
    o[0] = static_cast<uint8_t>(af); o[4] = static_cast<uint8_t>(ar); o[8] = static_cast<uint8_t>(vv); o[12] = static_cast<uint8_t>(hh);
  }
}

// ---- K0c: work-class ordering.  The cost of a read's walk is set by the number of places where it stops matching (each
// costs a fresh round of lookups and binary searches); lanes of a warp that walk reads of different cost finish at different
// times and the warp's instructions run with most lanes masked off.  The reads are therefore handed to the walk kernel
// grouped by a cheap cost class read off the k-mer masks: the number of maximal runs of dead windows (0 = every window
// may hit ... 6+), reads with a non-ACGT window last.  Counting sort: histogram, then scatter (order within a class is
// arbitrary; results do not depend on it).
static constexpr int kWorkClasses = 8;
__device__ __forceinline__ int workClass(const LaneParams& P, uint64_t r) {
  const int mate = r >= P.reads.n ? 1 : 0;
  const uint64_t ri = r - static_cast<uint64_t>(mate) * P.reads.n;
  uint32_t len = P.reads.fixedLen;
  if (P.reads.off[mate]) len = static_cast<uint32_t>(P.reads.off[mate][ri + 1] - P.reads.off[mate][ri]);
  const int k = static_cast<int>(P.ix.k);
  if (len > P.maxReadLen || static_cast<int>(len) < k) return 0;
  const int np = static_cast<int>(len) - k + 1;
  int runs = 0;
  bool bad = false;
  uint32_t carry = 0;  // dead bit of the previous window
  for (int j = 0; j * 32 < np; ++j) {
    const uint4 mk = __ldg(P.kmask + r * P.nw + j);
    const int nb = np - j * 32 < 32 ? np - j * 32 : 32;
    const uint32_t rng = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
    const uint32_t dead = (mk.z & (mk.w | (mk.x & mk.y))) & rng;
    bad |= ((~mk.z) & rng) != 0u;
    runs += __popc(dead & ~((dead << 1) | carry));
    carry = dead >> 31;
  }
  if (bad) return kWorkClasses - 1;
  return runs < kWorkClasses - 2 ? runs : kWorkClasses - 2;
}

__global__ void __launch_bounds__(256) work_class_hist_kernel(LaneParams P) {
  __shared__ uint32_t h[kWorkClasses];
  if (threadIdx.x < kWorkClasses) h[threadIdx.x] = 0;
  __syncthreads();
  for (uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < P.reads.numReads; r += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    atomicAdd(&h[workClass(P, r)], 1u);
  __syncthreads();
  if (threadIdx.x < kWorkClasses && h[threadIdx.x]) atomicAdd(P.classCtl + threadIdx.x, h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) work_class_scatter_kernel(LaneParams P) {
  __shared__ uint32_t cnt[kWorkClasses], base[kWorkClasses];
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t rounds = (P.reads.numReads + stride - 1) / stride;
  for (uint64_t it = 0; it < rounds; ++it) {
    const uint64_t r = it * stride + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (threadIdx.x < kWorkClasses) cnt[threadIdx.x] = 0;
    __syncthreads();
    int c = 0;
    uint32_t rank = 0;
    if (r < P.reads.numReads) { c = workClass(P, r); rank = atomicAdd(&cnt[c], 1u); }
    __syncthreads();
    if (threadIdx.x < kWorkClasses) {
      uint32_t off = 0;  // exclusive scan of the histogram
      for (int j = 0; j < static_cast<int>(threadIdx.x); ++j) off += P.classCtl[j];
      base[threadIdx.x] = off + (cnt[threadIdx.x] ? atomicAdd(P.classCtl + kWorkClasses + threadIdx.x, cnt[threadIdx.x]) : 0u);
    }
    __syncthreads();
    if (r < P.reads.numReads) P.order[base[c] + rank] = static_cast<uint32_t>(r);
    __syncthreads();
  }
}

// 32 bases (base q in bits 63:62) and 32 invalid bits (base q in bit 0) starting at forward position q >= 0.
template <int NT>
__device__ __forceinline__ void loadWin(const uint4* sm, int nw, int q, uint64_t& val, uint32_t& invw) {
  const int wi = q >> 5, sh = q & 31;
  const uint4 a = sm[wi * NT];
  uint4 b = make_uint4(0u, 0u, 0u, 0u);
  if (sh != 0 && wi + 1 < nw) b = sm[(wi + 1) * NT];
  const uint64_t p0 = (static_cast<uint64_t>(a.y) << 32) | a.x, p1 = (static_cast<uint64_t>(b.y) << 32) | b.x;
  val = sh ? ((p0 << (2 * sh)) | (p1 >> (64 - 2 * sh))) : p0;
  invw = __funnelshift_r(a.z, b.z, sh);
}

// k-mer at position p of a strand (Kmer::fromChars, include/Kmer.hpp:524-542): returns false and the PARTIAL word
// (codes before the first invalid base, the rest zero) when the window holds a non-ACGT base.
template <int NT>
__device__ __forceinline__ bool kmerAt(const uint4* sm, int nw, int L, int k, bool rc, int p, uint64_t& w) {
  const int q = rc ? L - k - p : p;
  uint64_t val;
  uint32_t invw;
  loadWin<NT>(sm, nw, q, val, invw);
  const uint64_t wf = val >> (64 - 2 * k);
  invw &= (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
  if (!rc) {
    w = wf;
    if (invw == 0u) return true;
    const int j = __ffs(invw) - 1;
    w = j == 0 ? 0ULL : (wf & ~((1ULL << (2 * (k - j))) - 1ULL));
    return false;
  }
  if (invw != 0u) {  // 'U' is a valid base on the reverse-complement strand
    uint32_t m = invw;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1u;
      if (((wf >> (2 * (k - 1 - b))) & 3ULL) == 3ULL) invw &= ~(1u << b);
    }
  }
  w = kmerRC(wf, k);
  if (invw == 0u) return true;
  const int j = (k - 1) - (31 - __clz(invw));  // first invalid base in reverse-complement order
  w = j == 0 ? 0ULL : (w & ~((1ULL << (2 * (k - j))) - 1ULL));
  return false;
}

// std::string::find_first_of("nN", from) on a strand.  On the reverse-complement strand every non-ACGTU base reads 'N'.
template <int NT>
__device__ __noinline__ int findNLane(const uint4* sm, int L, bool rc, int from) {
  if (from >= L) return INT_MAX;
  if (!rc) {
    int wi = from >> 5;
    uint32_t m = sm[wi * NT].w & (0xffffffffu << (from & 31));
    while (true) {
      if (m) return wi * 32 + __ffs(m) - 1;
      ++wi;
      if (wi * 32 >= L) return INT_MAX;
      m = sm[wi * NT].w;
    }
  }
  const int f0 = L - 1 - from;
  int wi = f0 >> 5;
  uint4 a = sm[wi * NT];
  uint32_t m = a.z & (0xffffffffu >> (31 - (f0 & 31)));
  while (true) {
    while (m) {
      const int b = 31 - __clz(m);
      m &= ~(1u << b);
      const uint64_t p0 = (static_cast<uint64_t>(a.y) << 32) | a.x;
      if (((p0 >> (62 - 2 * b)) & 3ULL) != 3ULL) return L - 1 - (wi * 32 + b);
    }
    if (--wi < 0) return INT_MAX;
    a = sm[wi * NT];
    m = a.z;
  }
}

__device__ __noinline__ uint64_t query8Bytes(const uint8_t* src, int L, bool rc, int p) {
  uint64_t out = 0;
  for (int j = 0; j < 8; ++j) {
    const int pos = p + j;
    if (pos < L) {
      const uint8_t ch = rc ? rcChar(__ldg(src + (L - 1 - pos))) : upperChar(__ldg(src + pos));
      out |= static_cast<uint64_t>(ch) << (8 * j);
    }
  }
  return out;
}

// 8 upper-cased characters of a strand at p .. p+7 (byte j = position p + j; positions >= L unspecified).
template <int NT>
__device__ __forceinline__ uint64_t query8(const uint4* sm, int nw, const uint8_t* src, int L, bool rc, int p) {
  const int q = rc ? L - 8 - p : p;
  const int qs = q < 0 ? 0 : q;
  uint64_t val;
  uint32_t invw;
  loadWin<NT>(sm, nw, qs, val, invw);
  uint32_t v = static_cast<uint32_t>(val >> 48);  // 8 bases, base qs in bits 15:14
  invw &= 0xffu;
  if (q < 0) { v >>= 2 * (-q); invw &= (1u << (8 + q)) - 1u; }
  if (invw != 0u) return query8Bytes(src, L, rc, p);  // a non-ACGT base in the window: exact characters from the original read
  if (rc) {  // reverse the eight 2-bit codes and complement them: the strand's own 8 bases, first base in bits 15:14
    v = __brev(v) >> 16;
    v = (((v & 0x5555u) << 1) | ((v >> 1) & 0x5555u)) ^ 0xffffu;
  }
  const uint32_t a = v >> 8, b = v & 0xffu;
  const uint32_t sa = ((a >> 6) & 3u) | (((a >> 4) & 3u) << 4) | (((a >> 2) & 3u) << 8) | ((a & 3u) << 12);
  const uint32_t sb = ((b >> 6) & 3u) | (((b >> 4) & 3u) << 4) | (((b >> 2) & 3u) << 8) | ((b & 3u) << 12);
  const uint32_t lo = __byte_perm(0x54474341u, 0u, sa);  // "ACGT"
  const uint32_t hi = __byte_perm(0x54474341u, 0u, sb);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// 32 text characters from `pos` as 2-bit codes (first character in bits 63:62) + their non-ACGT bits (first in bit 0).
__device__ __forceinline__ void textWin32(const DeviceIndex& ix, int64_t pos, uint64_t& tc, uint32_t& tinv) {
  const Words8 rec = ldg256(ix.text2 + (pos >> 5));
  const int s = static_cast<int>(pos & 31);
  const uint64_t c0 = (static_cast<uint64_t>(rec.v[1]) << 32) | rec.v[0], c1 = (static_cast<uint64_t>(rec.v[3]) << 32) | rec.v[2];
  tc = s ? ((c0 << (2 * s)) | (c1 >> (64 - 2 * s))) : c0;
  tinv = __funnelshift_r(rec.v[4], rec.v[5], s);
}

// 32 characters of a strand of the read from strand position qp, same format.
template <int NT>
__device__ __forceinline__ void queryWin32(const uint4* sm, int nw, int L, bool rc, int qp, uint64_t& qc, uint32_t& qinv) {
  if (!rc) { loadWin<NT>(sm, nw, qp, qc, qinv); return; }
  const int q = L - 32 - qp;  // strand positions qp .. qp+31 = forward positions q+31 .. q, complemented
  uint64_t val;
  uint32_t inv;
  loadWin<NT>(sm, nw, q < 0 ? 0 : q, val, inv);
  if (q < 0) { val >>= 2 * (-q); inv <<= -q; }
  qc = kmerRC(val, 32);
  qinv = __brev(inv);
}

// One suffix comparison of extendSearchNaive (include/SASearcher.hpp:160-176 and the two sentinel searches): query
// q[i] vs text[t + i] for i >= i0 while i < m and t + i < n; q[sentIdx] reads `sent` when sentIdx >= 0.  Returns the
// index at which the reference's inner loop stops; rel = -1 (query < text), +1 (query > text), 0 (ran off).
// Byte-exact form on the ASCII text, 8 characters per step: the fall-back of cmpSuffixPacked.
template <int NT>
__device__ __noinline__ int cmpSuffixAscii(const uint8_t* text, int64_t n, const uint4* sm, int nw, const uint8_t* src, int L, bool rc, int rb, int m,
                                           int32_t t, int i0, int sentIdx, uint32_t sent, int& rel) {
  const int64_t limL = n - static_cast<int64_t>(t);
  const int lim = limL < static_cast<int64_t>(m) ? static_cast<int>(limL) : m;
  rel = 0;
  if (i0 >= lim) return i0;
  int i = i0;
  const uint8_t* tp = text + t + i;
  const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(tp) & ~static_cast<uintptr_t>(7));
  const unsigned sh = static_cast<unsigned>(reinterpret_cast<uintptr_t>(tp) & 7) * 8;
  uint64_t lo = __ldg(a);
  for (;;) {
    const uint64_t hi = __ldg(a + 1);  // the text section is padded: reads past n are in bounds
    const uint64_t tw = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
    uint64_t qw = query8<NT>(sm, nw, src, L, rc, rb + i);
    const unsigned sd = static_cast<unsigned>(sentIdx - i);
    if (sd < 8u) qw = (qw & ~(0xffULL << (8 * sd))) | (static_cast<uint64_t>(sent) << (8 * sd));
    uint64_t x = tw ^ qw;
    const int nv = lim - i;
    if (nv < 8) x &= (1ULL << (8 * nv)) - 1ULL;
    if (x) {
      const int d = (__ffsll(static_cast<long long>(x)) - 1) >> 3;
      rel = ((qw >> (8 * d)) & 0xffULL) < ((tw >> (8 * d)) & 0xffULL) ? -1 : 1;
      return i + d;
    }
    i += 8;
    if (i >= lim) return lim;
    lo = hi;
    ++a;
  }
}

// The same comparison on the packed text, 32 characters per step: one 256-bit load of the text record, the query window
// from shared memory (reverse-complemented in registers for the rc strand), XOR + count-leading-zeros.  A, C, G, T order
// like their codes; a window with any other character before the first difference ('$' at a transcript end, N or IUPAC
// in the read) is handed to cmpSuffixAscii from the current index on.  The sentinels '#' / '{' order below / above
// every text character (checked when the index image is built).
template <int NT>
__device__ __forceinline__ int cmpSuffix(const LaneParams& P, const uint4* sm, uint32_t r, int L, bool rc, int rb, int m, int32_t t, int i0,
                                         int sentIdx, uint32_t sent, int& rel) {
  const int nw = static_cast<int>(P.nw);
  const int64_t limL = P.ix.n - static_cast<int64_t>(t);
  const int lim = limL < static_cast<int64_t>(m) ? static_cast<int>(limL) : m;
  rel = 0;
  if (i0 >= lim) return i0;
  int i = i0;
  bool exact = P.ix.text2 == nullptr;
  if (!exact) {
    const int cmpLim = (sentIdx >= 0 && sentIdx < lim) ? sentIdx : lim;  // ordinary characters live below cmpLim
    while (i < cmpLim) {
      uint64_t tc, qc;
      uint32_t tinv, qinv;
      textWin32(P.ix, static_cast<int64_t>(t) + i, tc, tinv);
      queryWin32<NT>(sm, nw, L, rc, rb + i, qc, qinv);
      const int left = cmpLim - i;
      const int nv = left < 32 ? left : 32;
      uint64_t x = tc ^ qc;
      uint32_t bad = tinv | qinv;
      if (nv < 32) { x &= ~0ULL << (64 - 2 * nv); bad &= (1u << nv) - 1u; }
      const int d = x ? (__clzll(static_cast<long long>(x)) >> 1) : nv;
      if (bad != 0u && (__ffs(bad) - 1) <= d) { exact = true; break; }
      if (d < nv) {
        rel = ((qc >> (62 - 2 * d)) & 3ULL) < ((tc >> (62 - 2 * d)) & 3ULL) ? -1 : 1;
        return i + d;
      }
      i += nv;
    }
    if (!exact) {
      if (cmpLim < lim) { rel = sent == '#' ? -1 : 1; return cmpLim; }  // the sentinel meets a text character
      return lim;
    }
  }
  const int mate = r >= P.reads.n ? 1 : 0;
  const uint64_t ri = r - static_cast<uint64_t>(mate) * P.reads.n;
  const uint8_t* src = P.reads.off[mate] ? P.reads.seq[mate] + P.reads.off[mate][ri] : P.reads.seq[mate] + ri * P.reads.fixedLen;
  return cmpSuffixAscii<NT>(P.ix.text, P.ix.n, sm, nw, src, L, rc, rb, m, t, i, sentIdx, sent, rel);
}

// k-mer and its reverse complement -> SA intervals; both table probes are in flight together.
// knownA / knownB: the filter already proved that key absent.
template <bool PHF>
__device__ __forceinline__ void hashFind2(const DeviceIndex& ix, uint64_t ka, uint64_t kb, bool knownA, bool knownB, int2& ra, int2& rb) {
  if (PHF) {  // -p index: the two BooPHF walks run one after the other (one inlined copy of the walk)
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const int2 res = (which ? knownB : knownA) ? make_int2(-1, -1) : phfFindImpl(ix, which ? kb : ka);
      if (which) rb = res; else ra = res;
    }
    return;
  }
  // two 16-byte slots share a 32-byte sector: both are examined per probe (the scan order is still slot by slot, so the
  // first empty slot ends the search exactly as in hashFind)
  uint64_t sa = mix64(ka) & ix.tableMask & ~1ULL, sb = mix64(kb) & ix.tableMask & ~1ULL;
  bool da = knownA, db = knownB;
  ra = make_int2(-1, -1); rb = make_int2(-1, -1);
  if (da && db) return;
  for (;;) {
    Words8 a, b;
    if (!da) a = ldg256(ix.table + sa);
    if (!db) b = ldg256(ix.table + sb);
    if (!da) {
      const uint64_t k0 = (static_cast<uint64_t>(a.v[1]) << 32) | a.v[0], k1 = (static_cast<uint64_t>(a.v[5]) << 32) | a.v[4];
      if (k0 == ka) { ra = make_int2(static_cast<int>(a.v[2]), static_cast<int>(a.v[3])); da = true; }
      else if (k0 == kEmptyKey) da = true;
      else if (k1 == ka) { ra = make_int2(static_cast<int>(a.v[6]), static_cast<int>(a.v[7])); da = true; }
      else if (k1 == kEmptyKey) da = true;
      else sa = (sa + 2) & ix.tableMask;
    }
    if (!db) {
      const uint64_t k0 = (static_cast<uint64_t>(b.v[1]) << 32) | b.v[0], k1 = (static_cast<uint64_t>(b.v[5]) << 32) | b.v[4];
      if (k0 == kb) { rb = make_int2(static_cast<int>(b.v[2]), static_cast<int>(b.v[3])); db = true; }
      else if (k0 == kEmptyKey) db = true;
      else if (k1 == kb) { rb = make_int2(static_cast<int>(b.v[6]), static_cast<int>(b.v[7])); db = true; }
      else if (k1 == kEmptyKey) db = true;
      else sb = (sb + 2) & ix.tableMask;
    }
    if (da && db) return;
  }
}

// SASearcher::lce (include/SASearcher.hpp:318-334) incl. its doubled start offset; --noSensitive only.
__device__ __noinline__ int lceLane(const int32_t* SA, const uint8_t* text, int64_t n, int64_t p1, int64_t p2, int startAt, int stopAt) {
  p1 = p1 < 0 ? 0 : (p1 >= n ? n - 1 : p1);
  p2 = p2 < 0 ? 0 : (p2 >= n ? n - 1 : p2);
  const int64_t o1 = static_cast<int64_t>(__ldg(SA + p1)) + startAt, o2 = static_cast<int64_t>(__ldg(SA + p2)) + startAt;
  const int64_t maxIndex = o1 > o2 ? o1 : o2;
  for (int len = startAt;; ++len) {
    if (!(maxIndex + len < n)) return len;
    const uint8_t a = __ldg(text + o1 + len), b = __ldg(text + o2 + len);
    if (a != b || a == '$' || len >= stopAt) return len;
  }
}

// kmerScores.emplace_back of spotCheck_ / the first-hit scan (include/SACollector.hpp:200-227,:417-430) as three bit
// sets over forward positions: tested, present in forward orientation, present in reverse-complement orientation.
__device__ __noinline__ void voteLane(uint32_t* votes, uint32_t vw, bool isRC, int p, int L, int k, bool mer, bool comp) {
  const int q = isRC ? (L - k - p) : p;
  const uint32_t word = static_cast<uint32_t>(q) >> 5, bit = 1u << (q & 31);
  if (votes[word] & bit) return;
  votes[word] |= bit;
  if (isRC ? comp : mer) votes[vw + word] |= bit;
  if (isRC ? mer : comp) votes[2 * vw + word] |= bit;
}

enum : int { LST_IDLE = 0, LST_SCAN, LST_WSTART, LST_MM, LST_EXTINIT, LST_EXT, LST_EXTDONE, LST_POSTMM, LST_WALKEND, LST_FINAL, LST_EXIT };
enum : uint32_t {
  LF_RC = 1u, LF_LAST = 2u, LF_DIDFWD = 4u, LF_FOUND = 8u, LF_NOMOREN = 16u, LF_FIRST = 32u, LF_SECOND = 64u, LF_OVF = 128u,
  LF_STAGE_SHIFT = 8u, LF_STAGE_MASK = 3u << 8,
};

#ifndef RAPMAP_LANE_REFILL
#define RAPMAP_LANE_REFILL 8   // fetch new reads once this many lanes of a warp are idle
#endif
#ifndef RAPMAP_LANE_SPIN
#define RAPMAP_LANE_SPIN 4    // double misses (by the filter) a lane may consume per trip
#endif
#ifndef RAPMAP_LANE_CHUNK
#define RAPMAP_LANE_CHUNK 256  // interval-arena records a warp reserves per atomic
#endif

// PHF: -p index (BooPHF walk instead of the dense table).  GENERAL: any strand-decision / skip mode; false = the default flags'
// coverage mode (disableNIP && strictCheck, include/SACollector.hpp:138) with the k-mer vote and NIP code compiled out: the
// walk is bound by instruction fetch (profiles/r02b: 6.2 of 15 stall cycles per issue are `no instruction`), so code the
// default configuration never runs is kept out of its kernel.
template <int NT, int MINB, bool PHF, bool GENERAL>
__global__ void __launch_bounds__(NT, MINB) sa_collect_lane_kernel(LaneParams P) {
  extern __shared__ __align__(16) uint8_t smemRaw[];
  uint4* smw = reinterpret_cast<uint4*>(smemRaw) + threadIdx.x;
  const uint4* sm = smw;
  const int lane = threadIdx.x & 31;
  const int k = static_cast<int>(P.ix.k);
  const int nw = static_cast<int>(P.nw);
  const DevOpts& o = P.opts;
  const bool disableNIP = GENERAL ? (o.disableNIP != 0) : true;
  const bool strictCheck = GENERAL ? (o.strictCheck != 0) : true;
  const bool useCov = disableNIP && strictCheck;  // include/SACollector.hpp:138
  const bool voteMode = strictCheck && !useCov;
  const uint32_t slot = blockIdx.x * NT + threadIdx.x;
  IntervalRec* scr = P.ivScratch + static_cast<size_t>(slot) * 2 * P.ivStride;
  uint32_t* votes = voteMode ? P.voteScratch + static_cast<size_t>(slot) * 3 * P.voteWords : nullptr;

  uint32_t chunkBase = 0, chunkLeft = 0;  // warp-uniform: reserved slice of the interval arena

  int st = LST_IDLE;
  uint32_t r = 0, flags = 0;
  int L = 0, rb = 0, lbIn = 0, ubIn = 0, l = 0, rr = 0, lcpLP = 0, lcpRP = 0, prevILow = 0, prevIHigh = 0;
  int mlen = 0, b0 = 0, b1 = 0, mQ = 0, pass = 0, guard = 0, prevMMPEnd = 0, nF = 0, nR = 0;
  uint32_t fwdHit = 0, rcHit = 0, fwdCov = 0, rcCov = 0;
  bool ready = false;   // a k-mer (w, at lookPos) is waiting for its lookups
  const bool mglob = P.masksInGlobal != 0u;
  auto maskWord = [&](int wq) -> uint4 { return mglob ? __ldg(P.kmask + static_cast<size_t>(r) * P.nw + wq) : sm[(nw + wq) * NT]; };
  uint64_t w = 0;
  int lookPos = 0;

  for (;;) {
    // ---------------- publish + refill: lanes whose read is finished wait (LST_FINAL) until enough lanes of the warp are
    // finished or idle, then all of them publish their interval lists in one go and the idle lanes take the next reads of
    // the batch.  (Publishing each read the moment it finishes ran this whole block once per read with one active lane:
    // a third of the kernel's instructions, profiles/r02b.)
    {
      const unsigned waiting = __ballot_sync(0xffffffffu, st == LST_IDLE || st == LST_FINAL);
      if (waiting) {
        const unsigned busy = __ballot_sync(0xffffffffu, st != LST_IDLE && st != LST_EXIT && st != LST_FINAL);
        if (__popc(waiting) >= RAPMAP_LANE_REFILL || busy == 0u) {
          const bool fin = st == LST_FINAL;
          int tot = 0;
          if (fin) {
            if (flags & LF_FOUND) {
              if (useCov) {  // strand decision by coverage (:283-288)
                if (fwdCov > rcCov + static_cast<uint32_t>(o.strictCheckSlack)) nR = 0;
                else if (rcCov > fwdCov + static_cast<uint32_t>(o.strictCheckSlack)) nF = 0;
              } else if (strictCheck) {  // k-mer "spot check" vote (:289-337)
                if (fwdHit > 0u && rcHit == 0u) nR = 0;
                else if (rcHit > 0u && fwdHit == 0u) nF = 0;
                else {
                  int fs = 0, rs = 0;
                  for (uint32_t j = 0; j < P.voteWords; ++j) {
                    const int tested = __popc(votes[j]);
                    fs += 2 * __popc(votes[P.voteWords + j]) - tested;
                    rs += 2 * __popc(votes[2 * P.voteWords + j]) - tested;
                  }
                  if (fs > rs) nR = 0;
                  else if (rs > fs) nF = 0;
                }
              }
              if (o.covReq > 0.0 && disableNIP) {  // :343-358
                if (nF > 0 && (static_cast<double>(fwdCov) / static_cast<double>(L)) < o.covReq) nF = 0;
                if (nR > 0 && (static_cast<double>(rcCov) / static_cast<double>(L)) < o.covReq) nR = 0;
              }
            } else { nF = 0; nR = 0; }
            tot = nF + nR;
          }
          const unsigned pm = __ballot_sync(0xffffffffu, fin && tot > 0);
          uint32_t off = 0;
          if (pm) {  // warp-aggregated reservation out of the warp's arena slice
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
              // A read whose records cannot be stored publishes an EMPTY summary (the later kernels of this attempt must not
              // follow ivOff into unwritten or out-of-range arena memory); the status bit makes the host grow and re-run.
              bool stored = false;
              if (flags & LF_OVF) atomicOr(P.status, kStatIvScratchFull);
              else if (static_cast<uint64_t>(off) + static_cast<uint64_t>(tot) > P.arenaCap) atomicOr(P.status, kStatIntervalArenaFull);
              else {
                for (int i = 0; i < nF; ++i) P.arena[off + i] = scr[i];
                for (int i = 0; i < nR; ++i) P.arena[off + nF + i] = scr[P.ivStride + i];
                stored = true;
              }
              if (!stored) { nF = 0; nR = 0; off = 0; }
            }
            ReadSummary s;
            s.ivOff = off; s.nFwd = static_cast<uint16_t>(nF); s.nRc = static_cast<uint16_t>(nR);
            s.readLen = static_cast<uint16_t>(L); s.found = (flags & LF_FOUND) ? 1 : 0; s.pad = 0;
            P.summ[r] = s;
            st = LST_IDLE;
          }

          const unsigned idle = __ballot_sync(0xffffffffu, st == LST_IDLE);
          const int leader = __ffs(idle) - 1;
          uint32_t base = 0;
          if (lane == leader) base = atomicAdd(P.readCursor, static_cast<uint32_t>(__popc(idle)));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (st == LST_IDLE) {
            const uint64_t slotNr = static_cast<uint64_t>(base) + __popc(idle & ((1u << lane) - 1u));
            if (slotNr >= P.reads.numReads) st = LST_EXIT;
            else {
              const uint64_t nr = __ldg(P.order + slotNr);
              r = static_cast<uint32_t>(nr);
              const int mate = nr >= P.reads.n ? 1 : 0;
              const uint64_t ri = nr - static_cast<uint64_t>(mate) * P.reads.n;
              uint32_t len = P.reads.fixedLen;
              if (P.reads.off[mate]) len = static_cast<uint32_t>(P.reads.off[mate][ri + 1] - P.reads.off[mate][ri]);
              if (len > P.maxReadLen) {
                atomicOr(P.status, kStatReadTooLong);
                ReadSummary s;
                s.ivOff = 0; s.nFwd = 0; s.nRc = 0; s.readLen = 0; s.found = 0; s.pad = 0;
                P.summ[r] = s;
              } else {
                const uint4* src = P.packed + static_cast<size_t>(nr) * P.nw;
                for (int j = 0; j < nw; ++j) smw[j * NT] = __ldg(src + j);
                if (!mglob) {
                  const uint4* msrc = P.kmask + static_cast<size_t>(nr) * P.nw;
                  for (int j = 0; j < nw; ++j) smw[(nw + j) * NT] = __ldg(msrc + j);
                }
                if (voteMode) for (uint32_t j = 0; j < 3 * P.voteWords; ++j) votes[j] = 0u;
                L = static_cast<int>(len);
                rb = 0; flags = 0; nF = 0; nR = 0; fwdHit = 0; rcHit = 0; fwdCov = 0; rcCov = 0; ready = false;
                st = LST_SCAN;
              }
            }
          }
        }
      }
      if (__ballot_sync(0xffffffffu, st != LST_EXIT) == 0u) break;
    }

    // ---------------- A: advance to the next k-mer that needs a lookup; k-mers the L2-resident filter proves absent in
    // both orientations are double misses and are consumed on the spot (up to RAPMAP_LANE_SPIN per trip): the ~31
    // windows that cover a sequencing error cost L2 round trips instead of trips through the whole state machine.
    bool knownM = false, knownC = false;  // filter verdict "absent" for the waiting k-mer / its reverse complement
    for (int spin = 0; spin < RAPMAP_LANE_SPIN; ++spin) {
      if (ready || !(st == LST_SCAN || st == LST_WSTART || st == LST_MM)) break;
      const bool rc = (flags & LF_RC) != 0u;
      // ---- dead positions (kmer_mask_kernel): a run of windows that are valid and either homopolymers or absent from the
      // index in both orientations is stepped over with a bit scan - exactly what the loop below would do one base at a time
      // (++rb, no counter moves; with the k-mer vote the double misses are marked tested).  In the first-hit scan a window
      // followed directly by an N is left to the loop (the "<=" of SACollector.hpp:178).
      if (st != LST_MM) {
        while (rb + k <= L) {
          const int q = rc ? L - k - rb : rb;
          const int wq = q >> 5, bq = q & 31;
          const uint4 mk = maskWord(wq);
          const uint32_t miss = mk.x & mk.y & mk.z;          // valid, absent in both orientations
          uint32_t dead = miss | (mk.w & mk.z);              // ... or a homopolymer
          if (st == LST_SCAN) {
            const uint32_t n0 = sm[wq * NT].w, n1 = (wq + 1 < nw) ? sm[(wq + 1) * NT].w : 0u;
            dead &= ~__funnelshift_r(n0, n1, static_cast<uint32_t>(k));  // N at q + k
          }
          int n;
          uint32_t run;
          if (!rc) {
            const uint32_t live = ~dead >> bq;
            n = live ? (__ffs(live) - 1) : (32 - bq);
            run = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) << bq;
          } else {
            const uint32_t live = ~dead << (31 - bq);
            n = live ? __clz(live) : (bq + 1);
            run = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) << (bq + 1 - n);
          }
          if (n == 0) break;
          if (voteMode && st != LST_SCAN) votes[wq] |= run & miss;
          rb += n;
          if (!rc ? (bq + n < 32) : (n <= bq)) break;  // the run ended inside this word
        }
      } else {
        const int lp = rb + mlen - (k - 1);
        const int q = rc ? L - k - lp : lp;
        const uint4 mk = maskWord(q >> 5);
        if ((mk.x & mk.y & mk.z) >> (q & 31) & 1u) {  // the k-mer after the interval is a double miss (:599-616, :671)
          if (voteMode) voteLane(votes, P.voteWords, rc, lp, L, k, false, false);
          st = LST_POSTMM;
          break;
        }
      }
      for (;;) {
        if (st == LST_MM) lookPos = rb + mlen - (k - 1);  // mismatching k-mer after an interval, :599-616
        else {
          if (rb + k > L) { st = (st == LST_SCAN) ? LST_FINAL : LST_WALKEND; break; }
          lookPos = rb;
          if (st == LST_SCAN && !(flags & LF_NOMOREN)) {  // first-hit scan, include/SACollector.hpp:167-237
            const int ip = findNLane<NT>(sm, L, false, rb);
            if (ip == INT_MAX) flags |= LF_NOMOREN;
            else if (ip <= rb + k) { rb = ip + 1; continue; }  // note <= (SACollector.hpp:178)
          }
        }
        const bool valid = kmerAt<NT>(sm, nw, L, k, rc, lookPos, w);
        if (st == LST_MM) {
          if (valid) ready = true; else st = LST_POSTMM;
          break;
        }
        if (st == LST_WSTART && !valid) {  // getSAHits_ loop head, :505-516
          const int ip = findNLane<NT>(sm, L, rc, rb);
          if (ip < rb + k) { rb = ip + 1; continue; }
        }
        if (isHomopolymer(w, k)) { ++rb; continue; }  // :520-536
        ready = true;
        break;
      }
      if (!ready || P.ix.filter == nullptr) break;
      {
        const int q = rc ? L - k - lookPos : lookPos;
        const uint4 mk = maskWord(q >> 5);
        if ((mk.z >> (q & 31)) & 1u) {  // valid window: the filter verdicts are in the masks (and it is not a double miss)
          const bool aF = (mk.x >> (q & 31)) & 1u, aR = (mk.y >> (q & 31)) & 1u;
          knownM = rc ? aR : aF;
          knownC = rc ? aF : aR;
          if (!(knownM && knownC)) break;
        }
      }
      if (!(knownM && knownC)) {
        uint64_t fw;
        uint32_t ma, mb;
        filterQuery(w, kmerRC(w, k), P.ix.filterShift, fw, ma, mb);
        const uint32_t fv = ldgKeep(P.ix.filter + fw);
        knownM = (fv & ma) != ma;
        knownC = (fv & mb) != mb;
      }
      if (!(knownM && knownC)) break;
      // double miss: no counter moves (strandHits / otherStrandHits, :541-546,:671); the walk steps one base (:673)
      ready = false; knownM = false; knownC = false;
      if (voteMode && st != LST_SCAN) voteLane(votes, P.voteWords, rc, lookPos, L, k, false, false);
      if (st == LST_MM) st = LST_POSTMM; else ++rb;
    }
    __syncwarp();  // all lanes meet before the table phase

    // ---------------- B: one pair of table lookups (k-mer, reverse complement)
    if (ready) {
      ready = false;
      int2 fm, fc;
      hashFind2<PHF>(P.ix, w, kmerRC(w, k), knownM, knownC, fm, fc);
      const bool hm = fm.x >= 0, hc = fc.x >= 0;
      const bool rc = (flags & LF_RC) != 0u;
      if (st == LST_SCAN) {
        if (hm) { ++fwdHit; if (hc) ++rcHit; }
        if (hc && !fwdHit) ++rcHit;
        if (fwdHit + rcHit > 0u) {
          if (voteMode) voteLane(votes, P.voteWords, false, lookPos, L, k, hm, hc);
          flags |= LF_FOUND;
          lbIn = fm.x; ubIn = fm.y;  // start interval of the forward walk (used only when fwdHit > 0)
          st = LST_WALKEND;
        } else ++rb;
      } else {
        if (rc) { rcHit += hm ? 1u : 0u; fwdHit += hc ? 1u : 0u; } else { fwdHit += hm ? 1u : 0u; rcHit += hc ? 1u : 0u; }
        if (voteMode) voteLane(votes, P.voteWords, rc, lookPos, L, k, hm, hc);
        if (st == LST_WSTART) {
          if (!hm) ++rb;  // :673
          else { lbIn = fm.x; ubIn = fm.y; st = LST_EXTINIT; }
        } else st = LST_POSTMM;
      }
    }

    // ---------------- C: set up extendSearchNaive for the interval [lbIn, ubIn)
    if (st == LST_EXTINIT) {
      lbIn = lbIn - 1 > 0 ? lbIn - 1 : 0;  // :553
      const bool firstAttempt = o.doChaining ? (rb == 0) : true;
      const int endPos = firstAttempt ? L : min(rb + k + o.maxMMPExtension, L);
      mQ = endPos - rb;
      flags = (flags & ~(LF_FIRST | LF_SECOND)) | (firstAttempt ? LF_FIRST : 0u);
      pass = (ubIn - lbIn == 2) ? 3 : 0;
      l = lbIn; rr = ubIn; lcpLP = k; lcpRP = k; prevILow = k; prevIHigh = k; mlen = k; guard = 0;
      st = LST_EXT;
    }

    // ---------------- D: one probe of the three binary searches (include/SASearcher.hpp:87-309)
    __syncwarp();
    if (st == LST_EXT) {
      int cc, i0, m, sentIdx = -1;
      uint32_t sent = 0;
      if (pass == 3) { cc = lbIn + 1; i0 = k; }
      else { cc = static_cast<int>((static_cast<int64_t>(l) + rr) >> 1); i0 = lcpLP < lcpRP ? lcpLP : lcpRP; }
      if (pass == 1 || pass == 2) { m = mlen + 1; sentIdx = m - 1; sent = pass == 1 ? '#' : '{'; }
      else m = mQ;
#ifdef RAPMAP_LDCG
      const int32_t t = __ldcg(P.ix.SA + cc);
#else
      const int32_t t = __ldg(P.ix.SA + cc);
#endif
      int rel;
      const int i = cmpSuffix<NT>(P, sm, r, L, (flags & LF_RC) != 0u, rb, m, t, i0, sentIdx, sent, rel);
      if (pass == 3) {  // :109-126
        b0 = lbIn + 1; b1 = ubIn; mlen = i;
        st = LST_EXTDONE;
      } else if (pass == 0) {  // :150-209
        bool plt = true;
        if (rel < 0) { if (i > prevIHigh) prevIHigh = i; }
        else if (rel > 0) { if (i > prevILow) prevILow = i; plt = false; }
        else if (i == m || static_cast<int64_t>(t) + i == P.ix.n) { if (i > prevIHigh) prevIHigh = i; }
        bool fin = false;
        if (plt) { if (cc == l + 1) fin = true; else { rr = cc; lcpRP = i; } }
        else { if (cc == rr - 1) fin = true; else { l = cc; lcpLP = i; } }
        if (fin) mlen = max(max(i, prevILow), prevIHigh);
        else if (++guard >= 80) fin = true;  // only a malformed index gets here; protects the GPU from a hang
        if (fin) { pass = 1; l = lbIn; rr = ubIn; lcpLP = k; lcpRP = k; guard = 0; }
      } else {  // :224-258 lower bound with '#', :270-304 upper bound with '{'
        bool fin = false;
        int bnd = ubIn;
        if (rel <= 0) { if (cc == l + 1) { bnd = cc; fin = true; } else { rr = cc; lcpRP = i; } }
        else { if (cc == rr - 1) { bnd = rr; fin = true; } else { l = cc; lcpLP = i; } }
        if (!fin && ++guard >= 80) { fin = true; bnd = ubIn; }
        if (fin) {
          if (pass == 1) { b0 = bnd; pass = 2; l = b0 - 1; rr = ubIn; lcpLP = k; lcpRP = k; guard = 0; }
          else { b1 = bnd; if (b0 == b1) ++b1; st = LST_EXTDONE; }  // :307
        }
      }
    }

    // ---------------- E: bookkeeping
    __syncwarp();
    if (st == LST_EXTDONE) {
      if (o.doChaining && (flags & LF_FIRST) && !(flags & LF_SECOND) && !(mlen >= L) && mlen >= k + o.maxMMPExtension) {  // :568-575
        mQ = min(rb + k + o.maxMMPExtension, L) - rb;
        flags |= LF_SECOND;
        pass = (ubIn - lbIn == 2) ? 3 : 0;
        l = lbIn; rr = ubIn; lcpLP = k; lcpRP = k; prevILow = k; prevIHigh = k; mlen = k; guard = 0;
        st = LST_EXT;
      } else {
        st = LST_POSTMM;
        const bool rc = (flags & LF_RC) != 0u;
        if (b1 > b0 && (b1 - b0) < o.maxInterval) {  // :578
          const int idx = rc ? nR : nF;
          if (static_cast<uint32_t>(idx) < P.ivStride) {
            IntervalRec rec;
            rec.begin = b0; rec.end = b1; rec.len = static_cast<uint16_t>(mlen); rec.qpos = static_cast<uint16_t>(rb);
            scr[(rc ? P.ivStride : 0u) + idx] = rec;
          } else flags |= LF_OVF;
          if (rc) ++nR; else ++nF;
          const int correction = prevMMPEnd > rb ? prevMMPEnd - rb : 0;
          if (rc) rcCov += static_cast<uint32_t>(mlen - correction); else fwdCov += static_cast<uint32_t>(mlen - correction);
          prevMMPEnd = rb + mlen;
          if (rb + mlen < L) st = LST_MM;
        }
        lbIn = b0; ubIn = b1;
      }
    }
    if (st == LST_POSTMM) {
      if ((flags & LF_LAST) || rb + mlen >= L) st = LST_WALKEND;  // :623, :630
      else {  // :634-657 next start: MMP skip, or the NIP skip when --noSensitive
        const int mismatchPos = rb + mlen;
        const int lce = disableNIP ? mlen : lceLane(P.ix.SA, P.ix.text, P.ix.n, lbIn, static_cast<int64_t>(ubIn) - 1, mlen, L - mismatchPos);
        const int skipMatch = mismatchPos - (k - 1), skipLCE = rb + lce - (k - 1);
        rb = skipMatch > skipLCE ? skipMatch : skipLCE;
        if (!disableNIP && lce > mlen && L > k) rb = rb < L - k ? rb : L - k;
        if (rb + k == L) flags |= LF_LAST;  // :663
        st = LST_WSTART;
      }
    }
    if (st == LST_WALKEND) {  // strand sequencing of SACollector::operator(), :247-281
      uint32_t stage = (flags & LF_STAGE_MASK) >> LF_STAGE_SHIFT;
      st = LST_FINAL;
      if (stage == 0u) {
        stage = 1u;
        if (fwdHit) {  // walk forward from the first hit with its interval (:247-254)
          flags = (flags | LF_DIDFWD) & ~(LF_RC | LF_LAST);
          prevMMPEnd = 0;
          st = LST_EXTINIT;
        }
      }
      if (st == LST_FINAL && stage == 1u) {
        stage = 2u;
        const bool checkRC = useCov ? (rcHit > 0u) : (rcHit >= fwdHit);  // :256
        if (checkRC) { flags = (flags | LF_RC) & ~LF_LAST; rb = 0; prevMMPEnd = 0; st = LST_WSTART; }
      }
      if (st == LST_FINAL && stage == 2u) {
        stage = 3u;
        const bool checkFwd = useCov ? (fwdHit > 0u) : (fwdHit >= rcHit);  // :270
        if (!(flags & LF_DIDFWD) && checkFwd) { flags &= ~(LF_RC | LF_LAST); rb = 0; prevMMPEnd = 0; st = LST_WSTART; }
      }
      flags = (flags & ~LF_STAGE_MASK) | (stage << LF_STAGE_SHIFT);
    }

  }
}

} // namespace rapmap_b200
