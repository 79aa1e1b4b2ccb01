// Host check of rapmap_b200/csrc/pack_swar.cuh against the per-base rule of pack_reads_kernel.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include "../../rapmap_b200/csrc/pack_swar.cuh"
using namespace rapmap_b200;
static void perBase(const uint8_t* s, uint64_t& codes, uint32_t& inv, uint32_t& nn) {
  codes = 0; inv = 0; nn = 0;
  for (int b = 0; b < 32; ++b) {
    const uint32_t ch = s[b], uc = ch & 0xDFu;
    const bool ok = uc == 'A' || uc == 'C' || uc == 'G' || uc == 'T';
    const uint32_t code = ok ? (((ch >> 1) ^ (ch >> 2)) & 3u) : (uc == 'U' ? 3u : (uc == 'N' ? 1u : 0u));
    codes |= static_cast<uint64_t>(code) << (62 - 2 * b);
    if (!ok) inv |= 1u << b;
    if (uc == 'N') nn |= 1u << b;
  }
}
// the switch-based originals (sel_aln.cuh nt4, kmer_utils.cuh rcChar)
static uint8_t nt4Ref(uint8_t c) {
  if (c < 4) return c;
  switch (c | 0x20) { case 'a': return 0; case 'c': return 1; case 'g': return 2; case 't': return 3; default: return 4; }
}
static uint8_t rcCharRef(uint8_t c) {
  switch (c | 0x20) {
    case 'a': return 'T'; case 'c': return 'G'; case 'g': return 'C'; case 't': return 'A';
    case 'u': return (c == 'U' || c == 'u') ? 'A' : 'N';
    default: return 'N';
  }
}
int main() {
  long badCode = 0;
  for (int v = 0; v < 256; ++v) {
    if (baseCode(static_cast<uint32_t>(v), false) != nt4Ref(static_cast<uint8_t>(v))) ++badCode;
    if (baseCode(static_cast<uint32_t>(v), true) != nt4Ref(rcCharRef(static_cast<uint8_t>(v)))) ++badCode;
  }
  std::printf("base codes: mismatches %ld\n", badCode);
  if (badCode) return 1;
  std::mt19937_64 rng(99);
  const char alpha[] = "ACGTacgtNnUuRYKMSWBDHV$#-*";
  long bad = 0, n = 0;
  uint8_t s[32];
  auto check = [&]() {
    uint32_t x[8];
    std::memcpy(x, s, 32);
    uint64_t c0, c1; uint32_t i0, i1, n0, n1;
    packBases32(x, c0, i0, n0);
    perBase(s, c1, i1, n1);
    ++n;
    if (c0 != c1 || i0 != i1 || n0 != n1) ++bad;
  };
  for (int v = 0; v < 256; ++v)          // every byte value at every position, among clean bases
    for (int pos = 0; pos < 32; ++pos) { std::memset(s, "ACGT"[pos & 3], 32); s[pos] = static_cast<uint8_t>(v); check(); }
  for (long it = 0; it < 300000; ++it) {
    const int mode = it % 3;
    for (auto& c : s) c = mode == 0 ? "ACGT"[rng() & 3] : (mode == 1 ? alpha[rng() % (sizeof(alpha) - 1)] : static_cast<uint8_t>(rng()));
    check();
  }
  std::printf("words %ld, mismatches %ld\n", n, bad);
  return bad ? 1 : 0;
}
