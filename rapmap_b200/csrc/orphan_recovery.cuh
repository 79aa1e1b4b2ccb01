// Orphan recovery (--recoverOrphans): selective_alignment::utils::recoverOrphans (reference
// include/SelectiveAlignmentUtils.hpp:35-257) and the one edlib call it makes,
//   edlibAlign(read, rlen, window, wlen, edlibNewAlignConfig(k = rlen / 4, EDLIB_MODE_HW, EDLIB_TASK_DISTANCE))
// (third-party edlib vendored in the reference at src/edlib.cpp:290): smallest edit distance of the whole read
// against any substring of the <= 1000-base window, -1 if it exceeds k, and the first 0-based end position reaching it.
//
// Edlib computes that with Myers' bit-vector algorithm in blocks of 64 query positions plus an Ukkonen band; the band
// only prunes, so the unbanded bit-vector recurrence below (Hyyro's formulation, one thread per anchor hit) returns the
// same (distance, first end position).  The function is host-callable so that tests can check it against the plain DP.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define RAPMAP_HD __host__ __device__
#else
#define RAPMAP_HD
#endif

namespace rapmap_b200 {

static constexpr int kOrphanMaxBlocks = 16;  // reads up to 1024 bases

// Query character i as edlib sees it: the raw read, or rapmap::utils::reverseRead of it (src/RapMapUtils.cpp:107-128).
RAPMAP_HD inline uint8_t orphanQueryChar(const uint8_t* read, int len, bool rc, int i) {
  if (!rc) return read[i];
  const uint8_t c = read[len - 1 - i];
  switch (c | 0x20) {
    case 'a': return 'T';
    case 'c': return 'G';
    case 'g': return 'C';
    case 't': return 'A';
    case 'u': return 'A';
    default: return 'N';
  }
}

// Returns the distance (or -1) and sets firstEnd.  m = query length (1 .. 64 * kOrphanMaxBlocks), n = window length.
RAPMAP_HD inline int semiGlobalMyers(const uint8_t* read, int m, bool rc, const uint8_t* win, int n, int k, int& firstEnd) {
  firstEnd = -1;
  if (m <= 0 || n <= 0 || m > 64 * kOrphanMaxBlocks) return -1;
  const int W = (m + 63) >> 6;
  uint64_t peq[4][kOrphanMaxBlocks], pv[kOrphanMaxBlocks], mv[kOrphanMaxBlocks];
  for (int b = 0; b < W; ++b) { peq[0][b] = peq[1][b] = peq[2][b] = peq[3][b] = 0; pv[b] = ~0ULL; mv[b] = 0; }
  for (int i = 0; i < m; ++i) {  // edlib compares bytes: only an exact 'A' / 'C' / 'G' / 'T' can equal a window base
    const uint8_t c = orphanQueryChar(read, m, rc, i);
    const int s = c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : (c == 'T' ? 3 : -1)));
    if (s >= 0) peq[s][i >> 6] |= 1ULL << (i & 63);
  }
  const uint64_t topBit = 1ULL << ((m - 1) & 63);  // last query position inside the last block
  int score = m, best = 0x7fffffff;
  for (int j = 0; j < n; ++j) {
    const uint8_t tc = win[j];
    const int s = tc == 'A' ? 0 : (tc == 'C' ? 1 : (tc == 'G' ? 2 : (tc == 'T' ? 3 : -1)));
    int hin = 0;  // free start in the window: the top row of the DP is all zeros
    for (int b = 0; b < W; ++b) {
      uint64_t eq;
      if (s >= 0) eq = peq[s][b];
      else {  // a window character outside ACGT: equality with the raw query bytes, computed on the spot
        eq = 0;
        for (int i = b * 64; i < m && i < b * 64 + 64; ++i)
          if (orphanQueryChar(read, m, rc, i) == tc) eq |= 1ULL << (i & 63);
      }
      const uint64_t pvb = pv[b], mvb = mv[b];
      const uint64_t xv = eq | mvb;
      if (hin < 0) eq |= 1ULL;
      const uint64_t xh = (((eq & pvb) + pvb) ^ pvb) | eq;
      uint64_t ph = mvb | ~(xh | pvb);
      uint64_t mh = pvb & xh;
      const uint64_t outBit = (b == W - 1) ? topBit : (1ULL << 63);
      int hout = 0;
      if (ph & outBit) hout = 1; else if (mh & outBit) hout = -1;
      ph <<= 1; mh <<= 1;
      if (hin < 0) mh |= 1ULL; else if (hin > 0) ph |= 1ULL;
      pv[b] = mh | ~(xv | ph);
      mv[b] = ph & xv;
      hin = hout;
    }
    score += hin;
    if (score < best) { best = score; firstEnd = j; }
  }
  if (best > k) { firstEnd = -1; return -1; }
  return best;
}

} // namespace rapmap_b200
