// ksw_extz2_sse41 in score-only mode (reference src/ksw2pp/ksw2_extz2_sse.c:18-304), TWO DP jobs of the same geometry
// per thread: the int8 SSE lanes of job 0 / job 1 live in the low / high 16-bit half of a 32-bit register, one
// register per band column, so every recurrence is ONE native sm_100a instruction for both jobs (VIADD.16x2,
// VIMNMX3.U16x2, VIMNMX.U16x2, VIADDMNMX.S16x2.RELU) instead of the 5-6 LOP3/IMAD/PRMT of an emulated byte-SIMD op, and
// nothing is shifted, masked or permuted per word: the band geometry (st, en, the 16-lane block rounding, the score
// window, the kcalloc restarts) is a function of (qlen, tlen, w) only and therefore shared by both jobs.
//
// Exactness.  The SSE code computes in wrapping int8 with a signed max, an unsigned max and an unsigned min.  With
// M = mat[0] + 2(q+e):  u, v stay in [0, M] for ANY stored x, y >= 0 (z = min(max3(s', a, b), M) >= vt1, ut), and as long as
// every stored x, y is < 32 and M + q <= 96 no int8 operation of the reference wraps, signed and unsigned compares
// agree, and 16-bit lanes compute the same numbers.  The kernel ORs all stored x, y together; a pair whose OR reaches
// 32 is handed to the byte-exact thread-per-job kernel (never seen in practice: x, y <= q for proper cells).
//
// H[] track.  The reference keeps an int32 H per column (H[t] += v[t] - qe inside the band, H[en0] = H[en0-1] + u[en0] - qe).
// Only H[st0] on the last query row (mqe) and H[en0] in the last target column (mte) are read.  Each cell keeps
// u(r,t) + v(r-1,t-1) == v(r,t) + u(r-1,t) by construction (both are z), so inside the band the column sums and row sums of
// the reference commute and the two read-outs follow from ONE running sum along each band edge:
//   A(r) = H[st0(r)] = A(r-1) + (st0 unchanged ? v[st0] : u[st0]) - qe,   B(r) = H[en0(r)] = B(r-1) + (en0 advanced ? u[en0] : v[en0]) - qe.
// The one geometry where the reference's own sums do NOT commute (a one-cell band that stays on the same column for two
// anti-diagonals: the row rule then reads the H of a column that has left the band) is detected and handed to the
// byte-exact kernel as well.  tests/cpp/ksw_pair_vs_ref.cpp checks this file (it is host-callable) against the
// reference's own SSE function on random and adversarial inputs.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define RAPMAP_HD __host__ __device__
#else
#define RAPMAP_HD
#endif

namespace rapmap_b200 {
namespace kswpair {

// ---- 16x2 primitives (native on sm_100a; plain C++ on the host for the test)
RAPMAP_HD inline uint32_t max3u(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  return __vimax3_u16x2(a, b, c);
#else
  auto m = [](uint32_t x, uint32_t y) { return x > y ? x : y; };
  return m(m(a & 0xffffu, b & 0xffffu), c & 0xffffu) | (m(m(a >> 16, b >> 16), c >> 16) << 16);
#endif
}
RAPMAP_HD inline uint32_t minu(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vminu2(a, b);
#else
  auto m = [](uint32_t x, uint32_t y) { return x < y ? x : y; };
  return m(a & 0xffffu, b & 0xffffu) | (m(a >> 16, b >> 16) << 16);
#endif
}
RAPMAP_HD inline uint32_t maxu(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vmaxu2(a, b);
#else
  auto m = [](uint32_t x, uint32_t y) { return x > y ? x : y; };
  return m(a & 0xffffu, b & 0xffffu) | (m(a >> 16, b >> 16) << 16);
#endif
}
RAPMAP_HD inline uint32_t maxs(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vmaxs2(a, b);
#else
  auto m = [](int16_t x, int16_t y) { return static_cast<uint32_t>(static_cast<uint16_t>(x > y ? x : y)); };
  return m(static_cast<int16_t>(a), static_cast<int16_t>(b)) | (m(static_cast<int16_t>(a >> 16), static_cast<int16_t>(b >> 16)) << 16);
#endif
}
RAPMAP_HD inline uint32_t add2(uint32_t a, uint32_t b) {  // wrapping per half
#ifdef __CUDA_ARCH__
  return __vadd2(a, b);
#else
  return ((a + b) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16);
#endif
}
RAPMAP_HD inline uint32_t addRelu(uint32_t a, uint32_t b) {  // max(a + b, 0) per signed half, for a >= 0 (so b is a neutral third operand: no zero register)
#ifdef __CUDA_ARCH__
  return __viaddmax_s16x2_relu(a, b, b);
#else
  auto f = [](int16_t x, int16_t y) { int32_t s = static_cast<int16_t>(static_cast<uint16_t>(x + y)); return static_cast<uint32_t>(s > 0 ? s : 0); };
  return f(static_cast<int16_t>(a), static_cast<int16_t>(b)) | (f(static_cast<int16_t>(a >> 16), static_cast<int16_t>(b >> 16)) << 16);
#endif
}
RAPMAP_HD inline uint32_t perm(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  uint32_t d;   // PTX prmt in its generic form: bit 3 of a selector nibble replicates the sign of the selected byte (__byte_perm masks that bit off)
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
#else
  const uint64_t v = (static_cast<uint64_t>(b) << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t n = (sel >> (4 * i)) & 15u;
    uint32_t byte = static_cast<uint32_t>((v >> (8 * (n & 7u))) & 0xffu);
    if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;   // PRMT's sign-replicate mode
    r |= byte << (8 * i);
  }
  return r;
#endif
}
// int8 wrap: sign-extend the low byte of each half
RAPMAP_HD inline uint32_t sx8(uint32_t a) { return perm(a, a, 0xA280u); }
// max(a, 0) per signed half
RAPMAP_HD inline uint32_t relu(uint32_t a) {
#ifdef __CUDA_ARCH__
  return __vimax_s16x2_relu(a, a);
#else
  return maxs(a, 0u);
#endif
}
RAPMAP_HD inline uint32_t funnelR(uint32_t lo, uint32_t hi, uint32_t sh) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, sh);
#else
  return static_cast<uint32_t>(((static_cast<uint64_t>(hi) << 32) | lo) >> (sh & 31));
#endif
}

// A[c] for a run-time c without moving the array to local memory: a uniform branch picks the group of eight registers,
// three levels of selects the register (a 32-way branch tree cost a quarter of the kernel's time in branch latency).
RAPMAP_HD inline uint32_t pick8(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7, int c) {
  const bool p0 = (c & 1) != 0, p1 = (c & 2) != 0, p2 = (c & 4) != 0;
  const uint32_t t0 = p0 ? a1 : a0, t1 = p0 ? a3 : a2, t2 = p0 ? a5 : a4, t3 = p0 ? a7 : a6;
  const uint32_t s0 = p1 ? t1 : t0, s1 = p1 ? t3 : t2;
  return p2 ? s1 : s0;
}
RAPMAP_HD inline uint32_t pick16(const uint32_t (&A)[32], int c) {
  return (c & 8) ? pick8(A[8], A[9], A[10], A[11], A[12], A[13], A[14], A[15], c) : pick8(A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7], c);
}
RAPMAP_HD inline uint32_t pick32(const uint32_t (&A)[32], int c) {
  switch (c >> 3) {
    case 0: return pick8(A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7], c);
    case 1: return pick8(A[8], A[9], A[10], A[11], A[12], A[13], A[14], A[15], c);
    case 2: return pick8(A[16], A[17], A[18], A[19], A[20], A[21], A[22], A[23], c);
    default: return pick8(A[24], A[25], A[26], A[27], A[28], A[29], A[30], A[31], c);
  }
}

// Scoring constants of a launch (KSW2Aligner's 5x5 matrix: match, mismatch, wildcard = 0; gap open q, gap extend e, band w).
struct Consts {
  int w, q, qe;
  uint32_t qe2x2;    // 2(q+e) in both halves: the score word of a kcalloc'ed lane (s = 0)
  uint32_t Mx2;      // max_sc = mat0 + 2(q+e)
  uint32_t qx2;      // q
  uint32_t Kq;       // q + 0x8000 per half
  uint32_t nqex2;    // -(q+e) per half (two's complement)
  uint32_t m0x4;     // mat0 + 2(q+e) in all four bytes
  uint32_t Kmul;     // mat0 - mat1
  uint32_t Dmul;     // matN - mat1
  bool ok;           // these scores can run on the pair kernel at all
};

RAPMAP_HD inline Consts makeConsts(int mat0, int mat1, int matN, int q, int e, int w) {
  Consts C{};
  C.w = w; C.q = q; C.qe = q + e;
  const int qe2 = 2 * (q + e), M = mat0 + qe2;
  int minSc = mat1 < matN ? mat1 : matN;
  minSc = minSc < mat0 ? minSc : mat0;
  C.ok = w >= 1 && w <= 15 && q >= 0 && e >= 0 && mat0 > 0 && mat1 <= 0 && matN >= mat1 && matN <= mat0 && -minSc <= qe2 && M + q <= 96 && mat0 - mat1 <= 255;
  C.qe2x2 = static_cast<uint32_t>(qe2) * 0x00010001u;
  C.Mx2 = static_cast<uint32_t>(M) * 0x00010001u;
  C.qx2 = static_cast<uint32_t>(q) * 0x00010001u;
  C.Kq = static_cast<uint32_t>(q + 0x8000) * 0x00010001u;
  C.nqex2 = static_cast<uint32_t>(static_cast<uint16_t>(-(q + e))) * 0x00010001u;
  C.m0x4 = static_cast<uint32_t>(M) * 0x01010101u;
  C.Kmul = static_cast<uint32_t>(mat0 - mat1);
  C.Dmul = static_cast<uint32_t>(matN - mat1);
  return C;
}

static constexpr uint32_t kNeg16x2 = 0x80008000u;   // "never set" marker of the two read-outs (KSW_NEG_INF)

// Strip layout (per thread, 16-bit cells = {job 0 code, job 1 code}, two cells per 32-bit word, word i at myW[i * NT]):
// cells [0, TL) target codes (nt4; zero beyond tlen), cells [TL, TL + qlen) the REVERSED query codes, zeros behind.
RAPMAP_HD inline int stripTL(int tlen) { return (tlen + 15) / 16 * 16 + 16; }
RAPMAP_HD inline int stripCells(int qlen, int tlen) { return stripTL(tlen) + qlen + 34; }

// Geometry the pair kernel takes (beyond Consts::ok): see the header comment.
RAPMAP_HD inline bool geomOk(const Consts& C, int qlen, int tlen, int maxCells) {
  if (!C.ok || qlen <= 0 || tlen < 64) return false;
  if (qlen >= tlen + C.w) return false;                               // a one-cell band could sit on the last column twice
  if (stripCells(qlen, tlen) > maxCells) return false;
  const int step = C.qe > static_cast<int>(C.Mx2 & 0xffffu) ? C.qe : static_cast<int>(C.Mx2 & 0xffffu);
  return (qlen + tlen + 2) * step < 32000;                           // the H sums fit the 16-bit halves
}

// The DP of one pair.  sc0 / sc1 = max(mqe, mte) of the two jobs (KSW_NEG_INF = -0x40000000 when neither was set);
// returns false when the pair has to be redone by the byte-exact kernel.
template <int NT>
RAPMAP_HD inline bool pairDP(const uint32_t* myW, int qlen, int tlen, const Consts& C, int32_t& sc0, int32_t& sc1) {
  const int TL = stripTL(tlen), w = C.w;
  uint32_t U[32], V[32], X[32], Y[32], S[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { U[c] = 0u; V[c] = 0u; X[c] = 0u; Y[c] = 0u; S[c] = C.qe2x2; }
  uint32_t A = 0u, B = 0u, mqe = kNeg16x2, mte = kNeg16x2, ovf = 0u;
  bool exact = true, moved = false, done = false;
  int curSt = 0, pst0 = 0, pen0 = 0, r = 0;
  const int R = qlen + tlen - 1;
  uint32_t xIn = 0u, vIn = 0u;   // OLD x / v of column st - 1 on the first anti-diagonal after the window moved
  // Outer loop: one pass per position of the 16-aligned window start; the state moves down 16 slots BETWEEN the inner loops
  // (a conditional move inside the anti-diagonal loop made the compiler re-copy all 160 state registers every iteration).
  while (!done) {
  for (; r < R; ++r) {
    int st0 = 0, en0 = tlen - 1;
    if (st0 < r - qlen + 1) st0 = r - qlen + 1;
    if (en0 > r) en0 = r;
    if (st0 < ((r - w + 1) >> 1)) st0 = (r - w + 1) >> 1;
    if (en0 > ((r + w) >> 1)) en0 = (r + w) >> 1;
    if (st0 > en0) { done = true; break; }
    const int st = st0 & ~15, en = ((en0 + 16) & ~15) - 1;
    if (st != curSt) break;   // the window start moves one block: shift below, then this anti-diagonal again
    uint32_t xPrev, vPrev;    // OLD x / v of column st - 1
    if (moved) { xPrev = xIn; vPrev = vIn; moved = false; }   // it was slot 15 before the move
    else if (st > 0) { xPrev = 0u; vPrev = 0u; }              // "not calculated; set to zeros" (:121)
    else { xPrev = 0u; vPrev = r ? C.qx2 : 0u; }              // :122
    const int cSt0 = st0 - st, cEn = en - st, cEn0 = en0 - st;
    if (en >= r) {  // the diagonal's first-row cell (:123), only during the first anti-diagonals
      const int cR = r - st;
      const uint32_t uR = r ? C.qx2 : 0u;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c == cR) { Y[c] = 0u; U[c] = uR; }
    }
    // ---- scores of the window [st0, st0 + 16) (:126-140), two columns x two jobs per word
    const uint32_t* tW = myW + static_cast<size_t>(st >> 1) * NT;
    const int qB = TL + (qlen - 1 - r) + st;
    const uint32_t* qW = myW + static_cast<size_t>(qB >> 1) * NT;
    const uint32_t qSh = static_cast<uint32_t>(qB & 1) * 16u;
    // (all 16 words, although only the 8-9 that overlap the window are used: skipping the others with a branch per word
    // was slower, 7.1 vs 6.6 ms - this kernel runs two warps per scheduler and pays for every branch)
    uint32_t qLo = qW[0];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t qHi = qW[static_cast<size_t>(k + 1) * NT];
      const uint32_t a1 = tW[static_cast<size_t>(k) * NT], a2 = funnelR(qLo, qHi, qSh);
      qLo = qHi;
      const uint32_t nf = ((a1 | a2) >> 2) & 0x01010101u;                          // code 4 on either side: wildcard
      const uint32_t f = ((((a1 ^ a2) + 0x07070707u) >> 3) & 0x01010101u) | nf;     // codes differ, or wildcard
      const uint32_t s4 = C.m0x4 - f * C.Kmul + nf * C.Dmul;                       // s + 2(q+e) per byte
      // column c is in the window iff cSt0 <= c < cSt0 + 16
      const bool in0 = k < 8 ? (2 * k >= cSt0) : (2 * k - 16 < cSt0);
      const bool in1 = k < 8 ? (2 * k + 1 >= cSt0) : (2 * k - 15 < cSt0);
      S[2 * k] = in0 ? perm(s4, s4, 0x9180u) : S[2 * k];          // bytes 0, 1 -> the two halves (the zero bytes are sign fills of bytes < 128)
      S[2 * k + 1] = in1 ? perm(s4, s4, 0xB3A2u) : S[2 * k + 1];  // bytes 2, 3
    }
    // ---- core loop (:147-164) over the computed lanes [st, en]
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (c == 16 && cEn < 16) break;
      const uint32_t xOld = X[c], vOld = V[c], uOld = U[c];
      const uint32_t aa = xPrev + vPrev;                          // checked below: never beyond 127, so it never wraps
      const uint32_t bb = Y[c] + uOld;                            // _mm_add_epi8: wraps in the out-of-band lanes (128 .. 127 + M), and the
      uint32_t z = max3u(S[c], aa, bb);                           // wrapped byte read as unsigned IS this number: max_epu8(z, b) as it stands
      z = minu(z, C.Mx2);                                         // min_epu8: z in [0, M]
      U[c] = z - vPrev;                                           // in [0, M] (z >= vt1, see the header comment)
      V[c] = z - uOld;
      const uint32_t c1 = (C.Kq - z) ^ kNeg16x2;                  // q - z per signed half
      X[c] = addRelu(aa, c1);
      Y[c] = relu(sx8(add2(bb, c1)));
      ovf = maxu(ovf, aa);
      xPrev = xOld; vPrev = vOld;
    }
    // ---- the two edge sums of the H[] track (see the header comment)
    if (r > 0) {
      const bool sameSt = st0 == pst0, sameEn = en0 == pen0;
      if (st0 == en0 && sameSt && sameEn) exact = false;
      const uint32_t dA = sameSt ? pick16(V, cSt0) : pick16(U, cSt0);
      const uint32_t dB = sameEn ? pick32(V, cEn0) : pick32(U, cEn0);
      A = add2(A, add2(dA, C.nqex2));
      B = add2(B, add2(dB, C.nqex2));
    } else {
      A = add2(V[0], add2(C.nqex2, C.nqex2));
      B = A;
    }
    if (en0 == tlen - 1) mte = maxs(mte, B);
    if (r - st0 == qlen - 1) mqe = maxs(mqe, A);
    pst0 = st0; pen0 = en0;
  }
  if (done || r >= R) break;
  xIn = X[15]; vIn = V[15];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    U[c] = U[c + 16]; V[c] = V[c + 16]; X[c] = X[c + 16]; Y[c] = Y[c + 16]; S[c] = S[c + 16];
    U[c + 16] = 0u; V[c + 16] = 0u; X[c + 16] = 0u; Y[c + 16] = 0u; S[c + 16] = C.qe2x2;
  }
  curSt += 16;
  moved = true;
  }
  const uint32_t best = maxs(mqe, mte);
  const int16_t b0 = static_cast<int16_t>(best & 0xffffu), b1 = static_cast<int16_t>(best >> 16);
  sc0 = b0 == static_cast<int16_t>(-32768) ? -0x40000000 : static_cast<int32_t>(b0);
  sc1 = b1 == static_cast<int16_t>(-32768) ? -0x40000000 : static_cast<int32_t>(b1);
  return exact && (ovf & 0xffffu) <= 127u && (ovf >> 16) <= 127u;   // no x + v ever wrapped
}

} // namespace kswpair
} // namespace rapmap_b200
