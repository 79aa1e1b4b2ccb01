// Deterministic synthetic data generator for tests and bench.
//
//   * transcriptome "GENCODE-like": genes -> shared exons -> isoforms (each exon kept w.p. 0.7),
//     so that suffix-array intervals are multi-transcript and hit lists need intersection.
//   * reads: 2 x readLen paired-end, fragment ~N(250,25), mate2 = reverse complement of the
//     fragment tail, random mate swap, per-base substitution / insertion / deletion / N noise.
//
// Everything is driven by a counter-based RNG (splitmix64 keyed by (seed, item index)), so any
// pair can be regenerated independently of batch / rank partitioning, on any machine.
//
// Built twice: as a CLI (`synth txome|reads ...`) and as a shared library (libsynth.so) whose
// extern "C" entry points fill caller-provided buffers (used by bench.py / tests via ctypes).
// This is data tooling, not part of the mapping path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

inline uint64_t splitmix(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ULL;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

struct Rng {
  uint64_t s;
  Rng(uint64_t seed, uint64_t stream, uint64_t idx) {
    uint64_t t = seed * 0xD1342543DE82EF95ULL + stream;
    uint64_t a = splitmix(t);
    t = a ^ (idx * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL);
    s = splitmix(t);
  }
  uint64_t next() { return splitmix(s); }
  // uniform in [0, n)
  uint64_t below(uint64_t n) { return static_cast<uint64_t>((static_cast<unsigned __int128>(next()) * n) >> 64); }
  // true with probability p_ppm / 1e6
  bool chance_ppm(uint32_t ppm) { return below(1000000) < ppm; }
};

const char BASES[4] = {'A', 'C', 'G', 'T'};

inline char comp(char c) {
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 'N';
  }
}

struct Txome {
  std::string text;                // concatenated transcripts, no separators
  std::vector<int64_t> off, len;   // per transcript
  std::vector<std::string> names;
};

// One gene: 4..16 exons of 80..400 nt; 1..10 distinct isoforms, each exon kept w.p. 0.7, len >= 300.
void gen_gene(uint64_t seed, uint64_t g, Txome& tx) {
  Rng r(seed, 1, g);
  int nEx = 4 + static_cast<int>(r.below(13));
  std::vector<std::string> exons(nEx);
  for (auto& e : exons) {
    int L = 80 + static_cast<int>(r.below(321));
    e.resize(L);
    for (int i = 0; i < L; ++i) e[i] = BASES[r.below(4)];
  }
  int nIso = 1 + static_cast<int>(r.below(10));
  std::vector<uint32_t> masks;
  int tries = 0;
  while (static_cast<int>(masks.size()) < nIso && tries < 200) {
    ++tries;
    uint32_t m = 0;
    int64_t L = 0;
    for (int e = 0; e < nEx; ++e)
      if (r.below(10) < 7) { m |= 1u << e; L += static_cast<int64_t>(exons[e].size()); }
    if (L < 300) continue;
    if (std::find(masks.begin(), masks.end(), m) != masks.end()) continue;
    masks.push_back(m);
  }
  int t = 0;
  for (uint32_t m : masks) {
    tx.off.push_back(static_cast<int64_t>(tx.text.size()));
    for (int e = 0; e < nEx; ++e)
      if (m >> e & 1) tx.text += exons[e];
    tx.len.push_back(static_cast<int64_t>(tx.text.size()) - tx.off.back());
    tx.names.push_back("G" + std::to_string(g) + ".T" + std::to_string(t++));
  }
}

// Optional repeat families: nFam elements of 300 nt, each inserted (<=2% divergence) into ~2% of
// transcripts; exercises large SA intervals and long hit lists.
void add_repeats(uint64_t seed, int nFam, Txome& tx) {
  if (nFam <= 0) return;
  std::vector<std::string> fam(nFam);
  Rng r(seed, 3, 0);
  for (auto& f : fam) {
    f.resize(300);
    for (auto& c : f) c = BASES[r.below(4)];
  }
  Txome out;
  for (size_t t = 0; t < tx.off.size(); ++t) {
    std::string s = tx.text.substr(tx.off[t], tx.len[t]);
    Rng q(seed, 4, t);
    if (q.below(100) < 2) {
      std::string el = fam[q.below(nFam)];
      for (auto& c : el)
        if (q.below(100) < 2) c = BASES[q.below(4)];
      size_t at = q.below(s.size() + 1);
      s.insert(at, el);
    }
    out.off.push_back(static_cast<int64_t>(out.text.size()));
    out.len.push_back(static_cast<int64_t>(s.size()));
    out.text += s;
  }
  out.names = tx.names;
  tx = std::move(out);
}

Txome gen_txome(uint64_t seed, int64_t genes, int repeatFamilies) {
  Txome tx;
  tx.text.reserve(static_cast<size_t>(genes) * 9500);
  for (int64_t g = 0; g < genes; ++g) gen_gene(seed, static_cast<uint64_t>(g), tx);
  add_repeats(seed, repeatFamilies, tx);
  return tx;
}

struct ErrModel {
  uint32_t sub_ppm{10000}, ins_ppm{300}, del_ppm{300}, n_ppm{1000};
};

// Emits exactly L bases of a noisy copy of src[0..srcLen) (src may run past the fragment into the
// transcript; random padding when exhausted).
void noisy_copy(Rng& r, const char* src, int64_t srcLen, bool rc, int L, const ErrModel& em, uint8_t* out) {
  int o = 0;
  int64_t i = 0;
  while (o < L) {
    char b;
    if (i < srcLen) {
      b = rc ? comp(src[srcLen - 1 - i]) : src[i];
    } else {
      b = BASES[r.below(4)];
    }
    uint64_t u = r.below(1000000);
    if (u < em.del_ppm) { ++i; continue; }
    u -= em.del_ppm;
    if (u < em.ins_ppm) { out[o++] = static_cast<uint8_t>(BASES[r.below(4)]); continue; }
    u -= em.ins_ppm;
    if (u < em.sub_ppm) {
      int k = 0;
      while (BASES[k] != b && k < 3) ++k;
      b = BASES[(k + 1 + r.below(3)) & 3];
    } else {
      u -= em.sub_ppm;
      if (u < em.n_ppm) b = 'N';
    }
    out[o++] = static_cast<uint8_t>(b);
    ++i;
  }
}

struct PairTruth { int64_t txp, pos, fragLen; int swapped; };

PairTruth gen_pair(uint64_t seed, int64_t idx, const char* text, const int64_t* off, const int64_t* len, int64_t ntxp,
                   int L, const ErrModel& em, uint8_t* m1, uint8_t* m2) {
  Rng r(seed, 2, static_cast<uint64_t>(idx));
  int64_t t = static_cast<int64_t>(r.below(static_cast<uint64_t>(ntxp)));
  int64_t tl = len[t];
  // Irwin-Hall(12) ~ N(6,1): integer-only so every platform agrees
  int64_t acc = 0;
  for (int i = 0; i < 12; ++i) acc += static_cast<int64_t>(r.below(65536));
  int64_t fl = 250 + ((acc - 6 * 65536) * 25) / 65536;
  int64_t lo = L + 10;
  if (fl < lo) fl = lo;
  if (fl > tl) fl = tl;
  int64_t pos = static_cast<int64_t>(r.below(static_cast<uint64_t>(tl - fl + 1)));
  const char* frag = text + off[t] + pos;
  int swapped = static_cast<int>(r.below(2));
  uint8_t* a = swapped ? m2 : m1;   // forward mate (fragment head)
  uint8_t* b = swapped ? m1 : m2;   // reverse-complement mate (fragment tail)
  // forward mate may read on past the fragment into the transcript
  noisy_copy(r, frag, tl - pos, false, L, em, a);
  // reverse mate: reverse complement of the fragment, reading towards the transcript start
  noisy_copy(r, text + off[t], pos + fl, true, L, em, b);
  return {t, pos, fl, swapped};
}

} // namespace

extern "C" {

struct SynthTxome {
  char* text; int64_t text_len; int64_t* off; int64_t* len; int64_t ntxp;
};

// Generates the transcriptome; caller frees with synth_txome_free.
SynthTxome* synth_txome_new(uint64_t seed, int64_t genes, int repeat_families) {
  Txome tx = gen_txome(seed, genes, repeat_families);
  auto* s = new SynthTxome();
  s->text_len = static_cast<int64_t>(tx.text.size());
  s->text = static_cast<char*>(malloc(tx.text.size() + 1));
  memcpy(s->text, tx.text.data(), tx.text.size());
  s->text[tx.text.size()] = 0;
  s->ntxp = static_cast<int64_t>(tx.off.size());
  s->off = static_cast<int64_t*>(malloc(sizeof(int64_t) * tx.off.size()));
  s->len = static_cast<int64_t*>(malloc(sizeof(int64_t) * tx.off.size()));
  memcpy(s->off, tx.off.data(), sizeof(int64_t) * tx.off.size());
  memcpy(s->len, tx.len.data(), sizeof(int64_t) * tx.len.size());
  return s;
}
void synth_txome_free(SynthTxome* s) {
  if (!s) return;
  free(s->text); free(s->off); free(s->len);
  delete s;
}
int64_t synth_txome_ntxp(const SynthTxome* s) { return s->ntxp; }
int64_t synth_txome_text_len(const SynthTxome* s) { return s->text_len; }

// Writes the transcriptome as FASTA (names G<g>.T<i> regenerated from the same seed).
int synth_txome_write_fasta(uint64_t seed, int64_t genes, int repeat_families, const char* path) {
  Txome tx = gen_txome(seed, genes, repeat_families);
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  for (size_t t = 0; t < tx.off.size(); ++t) {
    fprintf(f, ">%s\n", tx.names[t].c_str());
    fwrite(tx.text.data() + tx.off[t], 1, static_cast<size_t>(tx.len[t]), f);
    fputc('\n', f);
  }
  fclose(f);
  return 0;
}

// Fills out1/out2 (n * read_len bytes each, row-major) with pairs [first, first+n).
// truth (optional, 4 x int64 per pair): txp, pos, fragLen, swapped.
void synth_reads(const SynthTxome* tx, uint64_t seed, int64_t first, int64_t n, int read_len,
                 uint32_t sub_ppm, uint32_t ins_ppm, uint32_t del_ppm, uint32_t n_ppm,
                 uint8_t* out1, uint8_t* out2, int64_t* truth) {
  ErrModel em{sub_ppm, ins_ppm, del_ppm, n_ppm};
#pragma omp parallel for schedule(static, 4096)
  for (int64_t i = 0; i < n; ++i) {
    PairTruth pt = gen_pair(seed, first + i, tx->text, tx->off, tx->len, tx->ntxp, read_len, em,
                            out1 + i * read_len, out2 + i * read_len);
    if (truth) { truth[4 * i] = pt.txp; truth[4 * i + 1] = pt.pos; truth[4 * i + 2] = pt.fragLen; truth[4 * i + 3] = pt.swapped; }
  }
}

} // extern "C"

#ifdef SYNTH_MAIN
static const char* arg(int argc, char** argv, const char* k, const char* def) {
  for (int i = 1; i + 1 < argc; ++i)
    if (!strcmp(argv[i], k)) return argv[i + 1];
  return def;
}
int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr,
            "usage: synth txome --genes G --seed S [--repeats F] --out t.fasta\n"
            "       synth reads --genes G --seed S [--repeats F] --pairs N --rseed R [--first I] [--len L]\n"
            "                   [--sub ppm --ins ppm --del ppm --n ppm] --out1 r1.fastq --out2 r2.fastq\n");
    return 1;
  }
  uint64_t seed = strtoull(arg(argc, argv, "--seed", "12345"), nullptr, 10);
  int64_t genes = atoll(arg(argc, argv, "--genes", "100"));
  int reps = atoi(arg(argc, argv, "--repeats", "0"));
  if (!strcmp(argv[1], "txome")) {
    return synth_txome_write_fasta(seed, genes, reps, arg(argc, argv, "--out", "transcripts.fasta"));
  }
  if (!strcmp(argv[1], "reads")) {
    int64_t n = atoll(arg(argc, argv, "--pairs", "1000"));
    int64_t first = atoll(arg(argc, argv, "--first", "0"));
    uint64_t rseed = strtoull(arg(argc, argv, "--rseed", "54321"), nullptr, 10);
    int L = atoi(arg(argc, argv, "--len", "100"));
    uint32_t sub = atoi(arg(argc, argv, "--sub", "10000")), ins = atoi(arg(argc, argv, "--ins", "300")),
             del = atoi(arg(argc, argv, "--del", "300")), np = atoi(arg(argc, argv, "--n", "1000"));
    SynthTxome* tx = synth_txome_new(seed, genes, reps);
    std::vector<uint8_t> b1(static_cast<size_t>(n) * L), b2(static_cast<size_t>(n) * L);
    std::vector<int64_t> truth(static_cast<size_t>(n) * 4);
    synth_reads(tx, rseed, first, n, L, sub, ins, del, np, b1.data(), b2.data(), truth.data());
    FILE* f1 = fopen(arg(argc, argv, "--out1", "reads_1.fastq"), "w");
    FILE* f2 = fopen(arg(argc, argv, "--out2", "reads_2.fastq"), "w");
    if (!f1 || !f2) return 2;
    std::string qual(static_cast<size_t>(L), 'I');
    for (int64_t i = 0; i < n; ++i) {
      for (int m = 0; m < 2; ++m) {
        FILE* f = m ? f2 : f1;
        fprintf(f, "@r%lld:%lld:%lld:%lld/%d\n", static_cast<long long>(first + i), static_cast<long long>(truth[4 * i]),
                static_cast<long long>(truth[4 * i + 1]), static_cast<long long>(truth[4 * i + 2]), m + 1);
        fwrite((m ? b2.data() : b1.data()) + i * L, 1, static_cast<size_t>(L), f);
        fprintf(f, "\n+\n%s\n", qual.c_str());
      }
    }
    fclose(f1); fclose(f2);
    synth_txome_free(tx);
    return 0;
  }
  return 1;
}
#endif

// =================================================================================================
// hash.bin writer: lays a (k-mer -> SA interval) set out exactly as the reference's dense index file
// (sparsepp sparse_hash_map serialisation: magic 0x24687531, table_size, num_buckets as 4-byte
// big-endian words, one little-endian u32 occupancy bitmap per group of 32 buckets, then the records
// {u64 kmer, i32 begin, i32 end} in bucket order).  Bucket of a key = XXH64(key bytes, seed 0) & mask with
// triangular probing — the placement the reference's own find() walks.  Used by tools/build_index.py so
// that bench-scale indexes can be produced in seconds and still be read by the UNMODIFIED reference.
// =================================================================================================
namespace {
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
// XXH64 (public xxHash specification) specialised to an 8-byte input, seed 0.
inline uint64_t xxh64_u64(uint64_t v) {
  const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL,
                 P5 = 0x27D4EB2F165667C5ULL;
  uint64_t h = P5 + 8;
  uint64_t k1 = v * P2;
  k1 = rotl64(k1, 31);
  k1 *= P1;
  h ^= k1;
  h = rotl64(h, 27) * P1 + P4;
  h ^= h >> 33;
  h *= P2;
  h ^= h >> 29;
  h *= P3;
  h ^= h >> 32;
  return h;
}
void put_be32(FILE* f, uint32_t v) {
  unsigned char b[4] = {static_cast<unsigned char>(v >> 24), static_cast<unsigned char>(v >> 16), static_cast<unsigned char>(v >> 8),
                        static_cast<unsigned char>(v)};
  fwrite(b, 1, 4, f);
}
} // namespace

extern "C" int synth_write_dense_hash(const uint64_t* keys, const int32_t* begin, const int32_t* end, uint64_t n, const char* path) {
  uint64_t table = 32;
  while (table < 2 * n) table <<= 1;
  if (table >= 0xFFFFFFFFULL) return -2;
  const uint64_t mask = table - 1;
  std::vector<uint32_t> bitmap(table / 32, 0);
  std::vector<uint32_t> slotOf(n);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t b = xxh64_u64(keys[i]) & mask;
    uint64_t probes = 0;
    while (bitmap[b >> 5] >> (b & 31) & 1u) { ++probes; b = (b + probes) & mask; }
    bitmap[b >> 5] |= 1u << (b & 31);
    slotOf[i] = static_cast<uint32_t>(b);
  }
  // rank of each occupied bucket = position of its record in the file
  std::vector<uint32_t> groupBase(table / 32 + 1, 0);
  for (uint64_t g = 0; g < table / 32; ++g) groupBase[g + 1] = groupBase[g] + static_cast<uint32_t>(__builtin_popcount(bitmap[g]));
  struct Rec { uint64_t k; int32_t b, e; };
  std::vector<Rec> recs(n);
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t s = slotOf[i];
    uint32_t r = groupBase[s >> 5] + static_cast<uint32_t>(__builtin_popcount(bitmap[s >> 5] & ((1u << (s & 31)) - 1u)));
    recs[r] = {keys[i], begin[i], end[i]};
  }
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  put_be32(f, 0x24687531u);
  put_be32(f, static_cast<uint32_t>(table));
  put_be32(f, static_cast<uint32_t>(n));
  fwrite(bitmap.data(), 4, bitmap.size(), f);
  fwrite(recs.data(), sizeof(Rec), recs.size(), f);
  fclose(f);
  return 0;
}

// Transcript text with separators, as the reference indexer lays it out (src/RapMapSAIndexer.cpp:536-635):
// upper-case, poly-A tails of >= 10 clipped, '$' after every transcript.  Returns total length; fills
// out_text (caller-sized text_len + ntxp), starts (ntxp), complete lengths (ntxp).
extern "C" int64_t synth_txome_concat(const SynthTxome* s, char* out_text, int64_t* starts, uint32_t* complete_lens) {
  int64_t w = 0;
  for (int64_t t = 0; t < s->ntxp; ++t) {
    int64_t L = s->len[t];
    const char* p = s->text + s->off[t];
    complete_lens[t] = static_cast<uint32_t>(L);
    if (L > 10) {
      bool polyA = true;
      for (int i = 1; i <= 10; ++i) if (p[L - i] != 'A') { polyA = false; break; }
      if (polyA) { while (L > 0 && p[L - 1] == 'A') --L; }
    }
    if (L == 0) return -1;  // the reference would drop the entry; the generator never produces this
    starts[t] = w;
    memcpy(out_text + w, p, static_cast<size_t>(L));
    w += L;
    out_text[w++] = '$';
  }
  return w;
}
extern "C" void synth_txome_name(uint64_t seed, int64_t genes, int repeat_families, char* out, int64_t cap) {
  // names joined by '\n' (regenerated; cheap relative to indexing)
  Txome tx = gen_txome(seed, genes, repeat_families);
  int64_t w = 0;
  for (auto& nm : tx.names) {
    if (w + static_cast<int64_t>(nm.size()) + 1 >= cap) break;
    memcpy(out + w, nm.data(), nm.size());
    w += static_cast<int64_t>(nm.size());
    out[w++] = '\n';
  }
  out[w] = 0;
}

// =================================================================================================
// hash_info.bph / hash_info.val writer: the perfect-hash flavour of the index (`quasiindex -p`).
//   .bph = boomphf::mphf::save (reference include/BooPHF.hpp:1172-1197): gamma, #levels, last rank, #keys, then per level
//          {bits, words, bitset words, #rank samples, rank samples}, then the (here normally empty) final hash.
//   .val = FrugalBooMap::save (include/FrugalBooMap.hpp:185-213): data_ (interval start per MPHF slot, cereal vector<int32>),
//          lens_ (interval length, 255 => overflow_), overflow_ as a sparsepp table (spp_mix_32 placement, triangular probing).
// The MPHF is the published BBHash cascade as the reference builds it (BooPHF.hpp:885-968, :1233-1300, :1318-1366): level i
// has a bitset of ((ceil(gamma n) p^i) rounded up to 64) bits, p = 1 - ((gamma n - 1) / (gamma n))^(n - 1); every key that
// no earlier level placed hashes into it (xorshift128* sequence seeded by two hash64 values, fastrange64), positions hit
// exactly once stay set, the others are cleared and their keys go on; what reaches level 24 goes to an exact map.  The
// bitsets depend on the key SET only, so this writer's .bph equals the reference's own byte for byte whenever the final
// map is empty (tests/test_index_builder.py checks that, and that the unmodified reference maps with the files).
// =================================================================================================
namespace {
inline uint64_t boo_hash64(uint64_t key, uint64_t seed) {  // HashFunctors::hash64, BooPHF.hpp:394-407
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
// hash of `key` for level `lv`: h0, h1, then the xorshift128* sequence (XorshiftHashFunctors, BooPHF.hpp:460-499)
inline uint64_t boo_level_hash(uint64_t key, int lv) {
  uint64_t s0 = boo_hash64(key, 0xAAAAAAAA55555555ULL);
  if (lv == 0) return s0;
  uint64_t s1 = boo_hash64(key, 0x33333333CCCCCCCCULL);
  uint64_t h = s1;
  for (int i = 2; i <= lv; ++i) {
    uint64_t a = s0;
    const uint64_t b = s1;
    s0 = b;
    a ^= a << 23;
    s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
    h = s1 + b;
  }
  return h;
}
inline uint64_t fastrange64(uint64_t word, uint64_t p) { return static_cast<uint64_t>((static_cast<__uint128_t>(word) * static_cast<__uint128_t>(p)) >> 64); }
inline uint32_t spp_mix_32(uint32_t a) {
  a = a ^ (a >> 4);
  a = (a ^ 0xdeadbeef) + (a << 5);
  a = a ^ (a >> 11);
  return a;
}
} // namespace

extern "C" int synth_write_perfect_hash(const uint64_t* keys, const int32_t* begin, const int32_t* end, uint64_t n, const char* base) {
  if (n == 0) return -3;
  const double gamma = 2.0;
  const int nbLevels = 25;
  const double nD = static_cast<double>(n);
  const double proba = 1.0 - std::pow(((gamma * nD - 1) / (gamma * nD)), static_cast<double>(n - 1));
  const uint64_t hashDomain = static_cast<uint64_t>(std::ceil(nD * gamma));
  struct Level { uint64_t domain; std::vector<uint64_t> bits; std::vector<uint64_t> ranks; };
  std::vector<Level> lv(nbLevels);
  for (int i = 0; i < nbLevels; ++i) {
    uint64_t d = ((static_cast<uint64_t>(static_cast<double>(hashDomain) * std::pow(proba, i)) + 63) / 64) * 64;
    if (d == 0) d = 64;
    lv[i].domain = d;
    lv[i].bits.assign(1 + d / 64, 0);
  }
  // ---- the cascade
  std::vector<uint64_t> cur(keys, keys + n), nxt;
  std::vector<std::pair<uint64_t, uint64_t>> finalHash;
  for (int i = 0; i < nbLevels; ++i) {
    if (i == nbLevels - 1) {  // exact map for what is left (BooPHF.hpp:1098-1107)
      for (uint64_t j = 0; j < cur.size(); ++j) finalHash.emplace_back(cur[j], j);
      break;
    }
    Level& L = lv[i];
    std::vector<uint64_t> coll(L.bits.size(), 0);
    const int64_t m = static_cast<int64_t>(cur.size());
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m; ++j) {
      const uint64_t pos = fastrange64(boo_level_hash(cur[j], i), L.domain);
      const uint64_t bit = 1ULL << (pos & 63);
      const uint64_t old = __sync_fetch_and_or(&L.bits[pos >> 6], bit);
      if (old & bit) __sync_fetch_and_or(&coll[pos >> 6], bit);
    }
    for (size_t w = 0; w < L.bits.size(); ++w) L.bits[w] &= ~coll[w];
    // keys whose position was cleared go on to the next level
    nxt.clear();
#pragma omp parallel
    {
      std::vector<uint64_t> mine;
#pragma omp for schedule(static) nowait
      for (int64_t j = 0; j < m; ++j) {
        const uint64_t pos = fastrange64(boo_level_hash(cur[j], i), L.domain);
        if (!((L.bits[pos >> 6] >> (pos & 63)) & 1ULL)) mine.push_back(cur[j]);
      }
#pragma omp critical
      nxt.insert(nxt.end(), mine.begin(), mine.end());
    }
    cur.swap(nxt);
  }
  // ---- rank samples every 512 bits, offset by the keys of the earlier levels (bitVector::build_ranks, :739-753)
  uint64_t offset = 0;
  for (auto& L : lv) {
    for (size_t w = 0; w < L.bits.size(); ++w) {
      if ((w * 64) % 512 == 0) L.ranks.push_back(offset);
      offset += static_cast<uint64_t>(__builtin_popcountll(L.bits[w]));
    }
  }
  const uint64_t lastRank = offset;
  if (lastRank + finalHash.size() != n) return -4;  // duplicate keys in the input
  std::sort(finalHash.begin(), finalHash.end());
  // ---- slot of every key (mphf::lookup, :971-1009) -> data_ / lens_ in slot order
  std::vector<int32_t> data(n);
  std::vector<uint8_t> lens(n);
  std::vector<std::pair<int32_t, int32_t>> overflow;
  int bad = 0;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < static_cast<int64_t>(n); ++j) {
    const uint64_t key = keys[j];
    uint64_t slot = ~0ULL;
    int level = 0;
    uint64_t pos = 0;
    for (; level < nbLevels - 1; ++level) {
      pos = fastrange64(boo_level_hash(key, level), lv[level].domain);
      if ((lv[level].bits[pos >> 6] >> (pos & 63)) & 1ULL) break;
    }
    if (level == nbLevels - 1) {
      auto it = std::lower_bound(finalHash.begin(), finalHash.end(), std::make_pair(key, static_cast<uint64_t>(0)));
      if (it != finalHash.end() && it->first == key) slot = it->second + lastRank;
    } else {
      const Level& L = lv[level];
      const uint64_t wi = pos >> 6, block = pos >> 9;
      uint64_t r = L.ranks[block];
      for (uint64_t w = block * 8; w < wi; ++w) r += static_cast<uint64_t>(__builtin_popcountll(L.bits[w]));
      r += static_cast<uint64_t>(__builtin_popcountll(L.bits[wi] & ((1ULL << (pos & 63)) - 1ULL)));
      slot = r;
    }
    if (slot >= n) {
#pragma omp atomic write
      bad = 1;
      continue;
    }
    data[slot] = begin[j];
    const int32_t l = end[j] - begin[j];
    lens[slot] = l >= 255 ? 255 : static_cast<uint8_t>(l);
  }
  if (bad) return -5;
  for (uint64_t j = 0; j < n; ++j)
    if (end[j] - begin[j] >= 255) overflow.emplace_back(begin[j], end[j] - begin[j]);
  // ---- files
  {
    FILE* f = fopen((std::string(base) + ".bph").c_str(), "wb");
    if (!f) return -1;
    const int32_t nl = nbLevels;
    fwrite(&gamma, 8, 1, f); fwrite(&nl, 4, 1, f); fwrite(&lastRank, 8, 1, f); fwrite(&n, 8, 1, f);
    for (const auto& L : lv) {
      const uint64_t nchar = L.bits.size(), nr = L.ranks.size();
      fwrite(&L.domain, 8, 1, f); fwrite(&nchar, 8, 1, f); fwrite(L.bits.data(), 8, nchar, f);
      fwrite(&nr, 8, 1, f); fwrite(L.ranks.data(), 8, nr, f);
    }
    const uint64_t nf = finalHash.size();
    fwrite(&nf, 8, 1, f);
    for (const auto& kv : finalHash) { fwrite(&kv.first, 8, 1, f); fwrite(&kv.second, 8, 1, f); }
    fclose(f);
  }
  {
    FILE* f = fopen((std::string(base) + ".val").c_str(), "wb");
    if (!f) return -1;
    fwrite(&n, 8, 1, f); fwrite(data.data(), 4, n, f);
    fwrite(&n, 8, 1, f); fwrite(lens.data(), 1, n, f);
    // overflow_: sparse_hash_map<int32, int32>, hash spp_mix_32 (include/sparsepp/spp_utils.h:178-184)
    const uint64_t no = overflow.size();
    uint64_t table = 32;
    while (table * 4 < no * 5 + 4) table <<= 1;  // below sparsepp's 80 % occupancy
    const uint64_t mask = table - 1;
    std::vector<uint32_t> bitmap(table / 32, 0);
    std::vector<uint32_t> slotOf(no);
    for (uint64_t i = 0; i < no; ++i) {
      uint64_t b = spp_mix_32(static_cast<uint32_t>(overflow[i].first)) & mask, probes = 0;
      while (bitmap[b >> 5] >> (b & 31) & 1u) { ++probes; b = (b + probes) & mask; }
      bitmap[b >> 5] |= 1u << (b & 31);
      slotOf[i] = static_cast<uint32_t>(b);
    }
    std::vector<uint32_t> groupBase(table / 32 + 1, 0);
    for (uint64_t g = 0; g < table / 32; ++g) groupBase[g + 1] = groupBase[g] + static_cast<uint32_t>(__builtin_popcount(bitmap[g]));
    std::vector<std::pair<int32_t, int32_t>> recs(no);
    for (uint64_t i = 0; i < no; ++i) {
      const uint32_t s = slotOf[i];
      recs[groupBase[s >> 5] + static_cast<uint32_t>(__builtin_popcount(bitmap[s >> 5] & ((1u << (s & 31)) - 1u)))] = overflow[i];
    }
    put_be32(f, 0x24687531u);
    put_be32(f, static_cast<uint32_t>(table));
    put_be32(f, static_cast<uint32_t>(no));
    fwrite(bitmap.data(), 4, bitmap.size(), f);
    for (const auto& r : recs) { fwrite(&r.first, 4, 1, f); fwrite(&r.second, 4, 1, f); }
    fclose(f);
  }
  return 0;
}
