// Device image of a RapMapSAIndex and the primitive lookups the kernels share.
//
// HBM layout (one contiguous blob, sections 256-byte aligned, so the whole index is a single
// ncclBroadcast / cudaMemcpy):
//   ImageHeader | SA int32[n] | text u8[n + pad] | rank uint4[ceil(n/64)+1] | txpOffsets int32[T] |
//   txpLens int32[T] | hash table uint4[slots] | packed text TextRec[ceil(n/32)+1] | k-mer filter u32[F] | SA (tid, pos) uint2[n]
// * k-mer filter: a word-blocked Bloom filter over the indexed k-mers, keyed by the canonical k-mer with an orientation
//   marker (4 bits inside one 32-bit word, ~6 bits per key,
//   at most 64 MB so that it stays resident in the 126 MB L2, loads carry an L2 evict_last policy).  ~90 % of the
//   lookups of a read with sequencing errors are for k-mers that are not in the index (the 31 windows covering a
//   mismatch, both orientations); the filter answers ~92 % of those from L2 without touching the table in HBM.
//   No false negatives, so every lookup result is unchanged.  For -p indexes (no k-mer records on disk) the filter is
//   filled from the text: every key FrugalBooMap::find can return is verified against 31 text bases.
// * packed text: one 32-byte record per 32 text positions holding the 2-bit codes of the NEXT 64 positions and
//   their non-ACGT mask, so the 32-base window at ANY position p lies inside record p/32: one aligned 256-bit load
//   (one DRAM sector) feeds 32 character comparisons of extendSearchNaive.  The ASCII text stays for exact
//   fall-backs ('$', IUPAC) and for the selective-alignment windows.
// * SA (tid, pos): for every SA entry the transcript of the text position and the position inside it, precomputed at load
//   (8 n bytes of HBM): hit resolution expands an SA interval with ONE load per entry instead of the dependent chain
//   SA[i] -> rank record -> txpOffsets[tid].
// * rank: one 16-byte record per 64 text positions {bits_lo, bits_hi, #'$' before this word, 0}: a
//   transcript id is ONE 16-byte load + popcount (replaces rank9b::rank, reference src/rank9b.cpp:55-60,
//   which needs three loads; same value: number of set bits strictly before p).
// * hash table: open addressing, linear probing from the even slot of the key's 32-byte sector, 16-byte slots
//   {kmer_lo, kmer_hi, begin, end}, capacity a power of two >= 2 x #k-mers; a probe examines both slots of a sector.  Replaces the sparsepp
//   RegHashT<uint64_t, SAInterval> (reference include/RapMapUtils.hpp:65-67): same key -> [begin,end) map.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rapmap_b200 {

struct ImageHeader {
  uint64_t magic;          // 'RMB2IMG3'
  uint64_t totalBytes;
  uint64_t n;              // text / SA length
  uint64_t numTxp;
  uint64_t tableSlots;     // power of two
  uint64_t numKmers;
  uint32_t k;
  uint32_t hashKind;       // 0: dense open-addressing table, 1: BooPHF + FrugalBooMap arrays (-p index)
  uint64_t offSA, offText, offRank, offTxpOffsets, offTxpLens, offTable;
  uint64_t offFilter;      // 0: no filter
  uint64_t filterWords;    // power of two
  uint64_t offText2;       // 0: the text holds characters the packed compare cannot order (outside '$'..'z'); ASCII path only
  // -p index only
  uint32_t phfLevels;
  uint32_t phfTableDerived;  // 1: the hash-table section of this -p image answers every lookup (derived from FrugalBooMap::find at load time)
  uint64_t phfLastRank, phfNumData, phfNumFinal, phfNumOverflow;
  uint64_t offPhfLevels, offPhfBits, offPhfRanks, offPhfFinal, offPhfData, offPhfLens, offPhfOverflow;
  // host-readable trailer: transcript names, '\0'-terminated, in transcript order (a replica built from the image alone can
  // print SAM headers and records; the lengths are the txpLens section)
  uint64_t offNames, namesBytes;
  uint64_t offSaTidPos;    // uint2[n]: {transcript, position in it} of the text position SA[i] (hit resolution reads one record instead of SA -> rank -> txpOffsets)
};

// One level of the MPHF (boomphf::level, reference include/BooPHF.hpp:820-842): bitset of `domain` bits with a rank
// sample every 512 bits (bitVector::rank, :756-769).
struct PhfLevelDev {
  uint64_t domain;
  uint64_t bitsOff;    // first 64-bit word of this level in the shared bits array
  uint64_t ranksOff;   // first rank sample of this level
  uint64_t pad;
};
static constexpr uint64_t kImageMagic = 0x33474D4932424D52ULL;  // 'RMB2IMG3'

// 64 bases starting at text position 32 j: codes (first base in bits 63:62 of c0), non-ACGT mask (first base in bit 0 of inv0).
struct __align__(32) TextRec {
  uint32_t c0lo, c0hi, c1lo, c1hi, inv0, inv1, pad0, pad1;
};

struct DeviceIndex {
  const int32_t* SA;
  const uint8_t* text;
  const TextRec* text2;    // nullptr: no packed text (see ImageHeader::offText2)
  const uint32_t* filter;  // nullptr: no k-mer filter
  uint32_t filterShift;    // 64 - log2(filter words)
  const uint4* rank;
  const int32_t* txpOffsets;
  const int32_t* txpLens;
  const uint2* saTidPos;   // per SA entry: {transcript id, position inside the transcript}
  const uint4* table;
  uint64_t tableMask;
  int64_t n;
  uint32_t k;
  uint32_t numTxp;
  // perfect-hash flavour
  uint32_t hashKind, phfLevels, phfNumFinal, phfNumOverflow;
  uint64_t phfLastRank, phfNumData;
  const PhfLevelDev* phfLv;
  const uint64_t* phfBits;
  const uint64_t* phfRanks;
  const ulonglong2* phfFinal;    // sorted (key, value)
  const int32_t* phfData;        // FrugalBooMap::data_  : interval start per MPHF slot
  const uint8_t* phfLens;        // FrugalBooMap::lens_  : interval length, 255 => overflow
  const int2* phfOverflow;       // sorted (start, length)
};

static constexpr uint64_t kEmptyKey = ~0ULL;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// k-mer filter, keyed by the CANONICAL k-mer c = min(w, rc(w)): a read window asks about w and rc(w) at once, so one
// probe answers both.  From h = mix64(c): the word index comes from the top bits, three presence bits from bits 25..39, and
// an orientation marker from bits 40..44 (indexed k-mer equals its canonical form) or 45..49 (it is the reverse complement
// of it).  All five fields are disjoint from the word index for every filter size up to 2^14 ... 2^24 words.
//   w may be in the index   <=> the three presence bits and the marker of w's orientation are set
__host__ __device__ __forceinline__ void filterSlot(uint64_t h, uint32_t shift, uint64_t& word, uint32_t& presence, uint32_t& markCanon, uint32_t& markRc) {
  word = h >> shift;
  const uint32_t g = static_cast<uint32_t>(h >> 25);
  presence = (1u << (g & 31)) | (1u << ((g >> 5) & 31)) | (1u << ((g >> 10) & 31));
  markCanon = 1u << ((g >> 15) & 31);
  markRc = 1u << ((g >> 20) & 31);
}

// Reverse complement of a k-mer word (Kmer::getRC, reference include/Kmer.hpp:92-100), host and device.
__host__ __device__ __forceinline__ uint64_t kmerRevComp(uint64_t w, uint32_t k) {
  uint64_t x = ~w;  // complement: A<->T, C<->G is 3 - code
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
  x = (x >> 32) | (x << 32);
  return x >> (2 * (32 - k));
}

// Filter insertion of an indexed k-mer / query of a read k-mer w with reverse complement wr: the word and the two masks
// "w may be present" / "wr may be present" (a mask is satisfied when all its bits are set in the word).
__host__ __device__ __forceinline__ void filterInsertBits(uint64_t x, uint32_t k, uint32_t shift, uint64_t& word, uint32_t& bits) {
  const uint64_t r = kmerRevComp(x, k);
  const bool canon = x <= r;
  uint32_t p, mc, mr;
  filterSlot(mix64(canon ? x : r), shift, word, p, mc, mr);
  bits = p | (canon ? mc : mr);
}
__host__ __device__ __forceinline__ void filterQuery(uint64_t w, uint64_t wr, uint32_t shift, uint64_t& word, uint32_t& maskW, uint32_t& maskWr) {
  const bool canon = w <= wr;
  uint32_t p, mc, mr;
  filterSlot(mix64(canon ? w : wr), shift, word, p, mc, mr);
  maskW = p | (canon ? mc : mr);
  maskWr = p | (canon ? (w == wr ? mc : mr) : mc);  // a palindrome (even k only) is its own reverse complement
}

#ifdef __CUDACC__
// 32-bit load that asks L2 to keep the line (createpolicy evict_last): the k-mer filter is re-read ~50x per batch.
__device__ __forceinline__ uint32_t ldgKeep(const uint32_t* p) {
  uint32_t v;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// One aligned 256-bit load (LDG.E.256, sm_100): eight 32-bit words of one 32-byte DRAM sector.
struct __align__(32) Words8 { uint32_t v[8]; };
__device__ __forceinline__ Words8 ldg256(const void* p) {
  Words8 r;
#ifdef RAPMAP_LDCG
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif defined(RAPMAP_LDNA)
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

// boomphf hash of a key (HashFunctors::hash64, reference include/BooPHF.hpp:394-407)
__device__ __forceinline__ uint64_t phfHash64(uint64_t key, uint64_t seed) {
  uint64_t hash = seed;
  hash ^= (hash << 7) ^ key * (hash >> 3) ^ (~((hash << 11) + (key ^ (hash >> 5))));
  hash = (~hash) + (hash << 21);
  hash = hash ^ (hash >> 24);
  hash = (hash + (hash << 3)) + (hash << 8);
  hash = hash ^ (hash >> 14);
  hash = (hash + (hash << 2)) + (hash << 4);
  hash = hash ^ (hash >> 28);
  hash = hash + (hash << 31);
  return hash;
}

// The k-mer word FrugalBooMap::find builds from the text at a suffix (Kmer::fromChars; the text holds upper-case ACGT and '$').
__device__ __forceinline__ uint64_t phfTextWord(const DeviceIndex& ix, int64_t textInd) {
  uint64_t w = 0;
  for (uint32_t j = 0; j < ix.k; ++j) {
    const uint8_t ch = __ldg(ix.text + textInd + j);
    uint32_t cd;
    if (ch == 'A') cd = 0; else if (ch == 'C') cd = 1; else if (ch == 'G') cd = 2; else if (ch == 'T') cd = 3; else break;
    w |= static_cast<uint64_t>(cd) << (2 * (ix.k - 1 - j));
  }
  return w;
}

// FrugalBooMap::find (reference include/FrugalBooMap.hpp:149-167) over boomphf::mphf::lookup (include/BooPHF.hpp:971-1009,
// getLevel :1318-1351, xorshift128* next :493-499, fastrange64 :815-820, bitVector::rank :756-769): level walk -> rank ->
// data_[slot] -> SA -> verify the 31-mer in the text against the key -> length from lens_ / overflow_.
__device__ __forceinline__ int2 phfFindImpl(const DeviceIndex& ix, uint64_t key) {
  uint64_t s0 = 0, s1 = 0, h = 0, hashi = 0;
  uint32_t level = 0;
  const uint32_t last = ix.phfLevels - 1;
  for (uint32_t ii = 0; ii < last; ++ii) {
    if (ii == 0) { s0 = phfHash64(key, 0xAAAAAAAA55555555ULL); h = s0; }
    else if (ii == 1) { s1 = phfHash64(key, 0x33333333CCCCCCCCULL); h = s1; }
    else {
      uint64_t a = s0;
      const uint64_t b = s1;
      s0 = b;
      a ^= a << 23;
      s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
      h = s1 + b;
    }
    const PhfLevelDev lv = ix.phfLv[ii];
    hashi = __umul64hi(h, lv.domain);
    if ((__ldg(ix.phfBits + lv.bitsOff + (hashi >> 6)) >> (hashi & 63)) & 1ULL) break;
    ++level;
  }
  uint64_t slot;
  if (level == last) {  // _final_hash
    uint32_t lo = 0, hi = ix.phfNumFinal;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (ix.phfFinal[mid].x < key) lo = mid + 1; else hi = mid; }
    if (lo >= ix.phfNumFinal || ix.phfFinal[lo].x != key) return make_int2(-1, -1);
    slot = ix.phfFinal[lo].y + ix.phfLastRank;
  } else {
    const PhfLevelDev lv = ix.phfLv[level];
    const uint64_t wordIdx = hashi >> 6, block = hashi >> 9;
    uint64_t r = __ldg(ix.phfRanks + lv.ranksOff + block);
    for (uint64_t w = block * 8; w < wordIdx; ++w) r += __popcll(__ldg(ix.phfBits + lv.bitsOff + w));
    r += __popcll(__ldg(ix.phfBits + lv.bitsOff + wordIdx) & ((1ULL << (hashi & 63)) - 1ULL));
    slot = r;
  }
  if (slot >= ix.phfNumData) return make_int2(-1, -1);
  const int32_t ind = __ldg(ix.phfData + slot);
  if (phfTextWord(ix, __ldg(ix.SA + ind)) != key) return make_int2(-1, -1);
  int32_t len = __ldg(ix.phfLens + slot);
  if (len == 255) {
    uint32_t lo = 0, hi = ix.phfNumOverflow;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (ix.phfOverflow[mid].x < ind) lo = mid + 1; else hi = mid; }
    len = (lo < ix.phfNumOverflow && ix.phfOverflow[lo].x == ind) ? ix.phfOverflow[lo].y : 0;
  }
  return make_int2(ind, ind + len);
}

__device__ __noinline__ int2 phfFind(const DeviceIndex& ix, uint64_t key) { return phfFindImpl(ix, key); }

// k-mer -> SA interval; {-1,-1} when absent.  (RegHashT::find, reference include/SACollector.hpp:196,541)
__device__ __forceinline__ int2 hashFind(const DeviceIndex& ix, uint64_t key) {
  if (ix.hashKind) return phfFind(ix, key);
  uint64_t s = mix64(key) & ix.tableMask & ~1ULL;  // keys start at the even slot of their 32-byte sector
  while (true) {
    uint4 e = __ldg(ix.table + s);
    uint64_t kk = (static_cast<uint64_t>(e.y) << 32) | e.x;
    if (kk == key) return make_int2(static_cast<int>(e.z), static_cast<int>(e.w));
    if (kk == kEmptyKey) return make_int2(-1, -1);
    s = (s + 1) & ix.tableMask;
  }
}

// RapMapSAIndex::transcriptAtPosition (reference src/RapMapSAIndex.cpp:91-94): #'$' strictly before p.
__device__ __forceinline__ uint32_t transcriptAt(const DeviceIndex& ix, int64_t p) {
  uint4 r = __ldg(ix.rank + (p >> 6));
  uint64_t bits = (static_cast<uint64_t>(r.y) << 32) | r.x;
  uint64_t m = (1ULL << (p & 63)) - 1ULL;
  return r.z + __popcll(bits & m);
}
#endif

} // namespace rapmap_b200
