// 32 read bases (ASCII, eight little-endian 32-bit words) -> 2-bit codes + invalid mask + N mask, four bases per 32-bit
// operation.  Same result as the per-base rule of pack_reads_kernel: A C G T (either case) = 0..3; 'U' = 3, 'N' = 1, anything
// else 0, all three with the invalid bit; the N bit for 'N' / 'n'.  codes: base b in bits 63-2b..62-2b; masks: base b in bit b.
// Host-callable: tests/test_pack_swar_cpu.py checks it against the per-base rule on every byte value.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define RAPMAP_HD __host__ __device__
#else
#define RAPMAP_HD
#endif

namespace rapmap_b200 {

// 0x80 in every byte of v that is zero (exact: no borrow crosses a byte)
RAPMAP_HD inline uint32_t zeroBytes(uint32_t v) { return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u; }

RAPMAP_HD inline void packBases32(const uint32_t (&x)[8], uint64_t& codes, uint32_t& inv, uint32_t& nn) {
  codes = 0; inv = 0; nn = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < 8; ++j) {
    const uint32_t u = x[j] & 0xDFDFDFDFu;
    const uint32_t ok = zeroBytes(u ^ 0x41414141u) | zeroBytes(u ^ 0x43434343u) | zeroBytes(u ^ 0x47474747u) | zeroBytes(u ^ 0x54545454u);
    uint32_t code = ((x[j] >> 1) ^ (x[j] >> 2)) & 0x03030303u;
    if (ok != 0x80808080u) {  // a base that is not A/C/G/T
      const uint32_t isU = zeroBytes(u ^ 0x55555555u) >> 7, isN = zeroBytes(u ^ 0x4E4E4E4Eu) >> 7, bad = (~ok & 0x80808080u) >> 7;
      code = (code & ((ok >> 7) * 3u)) | (isU * 3u) | isN;
      inv |= (((bad * 0x00204081u) >> 21) & 0xFu) << (4 * j);   // byte i of the word -> bit i (the partial products fall on distinct bits)
      nn |= (((isN * 0x00204081u) >> 21) & 0xFu) << (4 * j);
    }
    const uint32_t c8 = (code * 0x40100401u) >> 24;              // the four 2-bit codes, first base on top
    codes |= static_cast<uint64_t>(c8) << (56 - 8 * j);
  }
}

// ksw2's code of one base (seq_nt4_table_loc, reference src/ksw2pp/KSW2Aligner.cpp:61-72: A C G T either case = 0..3, the raw
// bytes 0..3 themselves, anything else 4) without a branch; comp: the code of rapmap::utils::reverseRead's output for this
// byte (A<->T, C<->G, U -> A, anything else N = 4).  Used by the ksw2 pair kernel's strip fill.
RAPMAP_HD inline uint32_t baseCode(uint32_t ch, bool comp) {
  const uint32_t d = (ch & 0xDFu) - 'A';                                  // A C G T -> 0 2 6 19
  const bool acgt = d < 20u && ((0x80045u >> d) & 1u);
  const uint32_t c2 = ((ch >> 1) ^ (ch >> 2)) & 3u;
  if (!comp) return acgt ? c2 : (ch < 4u ? ch : 4u);
  return acgt ? 3u - c2 : ((ch & 0xDFu) == 'U' ? 0u : 4u);
}

} // namespace rapmap_b200
