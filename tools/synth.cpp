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
