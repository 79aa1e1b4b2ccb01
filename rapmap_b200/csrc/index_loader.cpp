#include "index_loader.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rapmap_b200 {

namespace {

struct File {
  FILE* f{nullptr};
  explicit File(const std::string& p) : f(std::fopen(p.c_str(), "rb")) {}
  ~File() { if (f) std::fclose(f); }
  bool ok() const { return f != nullptr; }
  bool read(void* p, size_t n) { return n == 0 || std::fread(p, 1, n, f) == n; }
  template <class T> bool get(T& v) { return read(&v, sizeof(T)); }
  uint64_t size() {
    long cur = std::ftell(f);
    std::fseek(f, 0, SEEK_END);
    long e = std::ftell(f);
    std::fseek(f, cur, SEEK_SET);
    return static_cast<uint64_t>(e);
  }
};

// header.json is cereal JSON ({"value0": {...}}, src/RapMapSAIndexer.cpp:791-818, include/IndexHeader.hpp:46-57);
// only three scalar fields matter here.
bool jsonField(const std::string& t, const char* key, std::string& out) {
  std::string pat = std::string("\"") + key + "\"";
  size_t p = t.find(pat);
  if (p == std::string::npos) return false;
  p = t.find(':', p + pat.size());
  if (p == std::string::npos) return false;
  ++p;
  while (p < t.size() && (t[p] == ' ' || t[p] == '\t' || t[p] == '\n')) ++p;
  size_t e = t.find_first_of(",}\n", p);
  out = t.substr(p, e - p);
  return true;
}

uint32_t be32(const unsigned char* b) {
  return (uint32_t(b[0]) << 24) | (uint32_t(b[1]) << 16) | (uint32_t(b[2]) << 8) | uint32_t(b[3]);
}

// sparsepp writes table metadata as "4 bytes big-endian, or 0xFFFFFFFF + 8 bytes big-endian"
bool read32or64(File& f, uint64_t& v, uint64_t& consumed) {
  unsigned char b[8];
  if (!f.read(b, 4)) return false;
  consumed += 4;
  uint32_t x = be32(b);
  if (x != 0xFFFFFFFFu) { v = x; return true; }
  if (!f.read(b, 8)) return false;
  consumed += 8;
  v = (uint64_t(be32(b)) << 32) | be32(b + 4);
  return true;
}

inline int code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

} // namespace

bool HostIndex::load(const std::string& dirIn, std::string& err) {
  std::string dir = dirIn;
  if (!dir.empty() && dir.back() != '/') dir += '/';

  {
    std::ifstream h(dir + "header.json");
    if (!h) { err = "cannot open " + dir + "header.json"; return false; }
    std::stringstream ss;
    ss << h.rdbuf();
    std::string t = ss.str(), v;
    if (!jsonField(t, "KmerLen", v)) { err = "header.json: no KmerLen"; return false; }
    k = static_cast<uint32_t>(std::strtoul(v.c_str(), nullptr, 10));
    bigSA = jsonField(t, "BigSA", v) && v.find("true") != std::string::npos;
    perfectHash = jsonField(t, "PerfectHash", v) && v.find("true") != std::string::npos;
    if (jsonField(t, "IndexVersion", v) && v.find("q5") == std::string::npos) {
      err = "header.json: unsupported IndexVersion " + v + " (expected q5)";
      return false;
    }
  }
  if (k == 0 || k > 31) { err = "unsupported k-mer length"; return false; }
  // BigSA indexes (IndexT = int64_t: src/RapMapSAIndexer.cpp:86-220, instantiations src/RapMapSAIndex.cpp:178-185) store 8-byte
  // suffix-array entries, transcript offsets, hash intervals and FrugalBooMap starts.  The device image keeps 32-bit positions,
  // so such an index is read and NARROWED when its text has fewer than 2^31 positions; a longer text is refused, never truncated.
  auto tooBig = [&]() {
    err = "BigSA index with a text of 2^31 or more positions: the device index holds 32-bit positions (SURVEY.md section 8 f4)";
    unsupported = true;
    return false;
  };

  {
    File f(dir + "sa.bin");
    if (!f.ok()) { err = "cannot open sa.bin"; return false; }
    uint64_t n = 0;
    if (!f.get(n)) { err = "sa.bin: truncated"; return false; }
    const uint64_t esz = bigSA ? 8 : 4;
    if (n > (f.size() - 8) / esz) { err = "sa.bin: element count exceeds the file size"; return false; }
    if (n >= (1ull << 31)) { if (bigSA) return tooBig(); err = "sa.bin: more than 2^31 suffixes in a 32-bit index"; return false; }
    SA.resize(n);
    if (!bigSA) {
      if (!f.read(SA.data(), n * sizeof(int32_t))) { err = "sa.bin: truncated"; return false; }
    } else {
      std::vector<int64_t> wide(std::min<uint64_t>(n, 1 << 20));
      for (uint64_t at = 0; at < n; at += wide.size()) {
        const uint64_t cnt = std::min<uint64_t>(wide.size(), n - at);
        if (!f.read(wide.data(), cnt * 8)) { err = "sa.bin: truncated"; return false; }
        for (uint64_t i = 0; i < cnt; ++i) {
          if (wide[i] < 0 || wide[i] >= static_cast<int64_t>(n)) { err = "sa.bin: suffix array entry outside the text"; return false; }
          SA[at + i] = static_cast<int32_t>(wide[i]);
        }
      }
    }
  }
  {
    File f(dir + "txpInfo.bin");
    if (!f.ok()) { err = "cannot open txpInfo.bin"; return false; }
    uint64_t n = 0;
    const uint64_t fsz = f.size();
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    if (n > fsz / 8) { err = "txpInfo.bin: transcript count exceeds the file size"; return false; }
    txpNames.resize(n);
    for (auto& s : txpNames) {
      uint64_t l = 0;
      if (!f.get(l)) { err = "txpInfo.bin: truncated"; return false; }
      if (l > fsz) { err = "txpInfo.bin: name length exceeds the file size"; return false; }
      s.resize(l);
      if (!f.read(&s[0], l)) { err = "txpInfo.bin: truncated"; return false; }
    }
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    if (n > fsz / (bigSA ? 8 : 4)) { err = "txpInfo.bin: offset count exceeds the file size"; return false; }
    txpOffsets.resize(n);
    if (!bigSA) {
      if (!f.read(txpOffsets.data(), n * sizeof(int32_t))) { err = "txpInfo.bin: truncated"; return false; }
    } else {
      std::vector<int64_t> wide(n);
      if (!f.read(wide.data(), n * 8)) { err = "txpInfo.bin: truncated"; return false; }
      for (uint64_t i = 0; i < n; ++i) {
        if (wide[i] < 0 || wide[i] >= (1ll << 31)) return tooBig();
        txpOffsets[i] = static_cast<int32_t>(wide[i]);
      }
    }
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    if (n > fsz) { err = "txpInfo.bin: text length exceeds the file size"; return false; }
    text.resize(n);
    if (!f.read(&text[0], n)) { err = "txpInfo.bin: truncated"; return false; }
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    if (n > fsz / 4) { err = "txpInfo.bin: length count exceeds the file size"; return false; }
    txpCompleteLens.resize(n);
    if (!f.read(txpCompleteLens.data(), n * sizeof(uint32_t))) { err = "txpInfo.bin: truncated"; return false; }
  }
  if (txpOffsets.empty() || txpOffsets.size() != txpNames.size()) { err = "txpInfo.bin: inconsistent transcript tables"; return false; }
  if (SA.size() != text.size()) { err = "sa.bin / txpInfo.bin: suffix array and text lengths differ"; return false; }
  {
    File f(dir + "rsd.bin");
    if (!f.ok()) { err = "cannot open rsd.bin"; return false; }
    if (!f.get(numBits)) { err = "rsd.bin: truncated"; return false; }
    if (numBits != SA.size()) { err = "rsd.bin: bit count differs from the text length"; return false; }
    uint64_t nbytes = (numBits + 7) / 8;
    rsdBits.assign((numBits + 63) / 64 + 1, 0);
    if (!f.read(rsdBits.data(), nbytes)) { err = "rsd.bin: truncated"; return false; }
  }
  {  // cheap range checks: a malformed index must fail here, not as out-of-bounds reads on the device
    const int64_t n = static_cast<int64_t>(SA.size());
    for (size_t i = 0; i < txpOffsets.size(); ++i) {
      if (txpOffsets[i] < 0 || txpOffsets[i] >= n || (i > 0 && txpOffsets[i] <= txpOffsets[i - 1])) { err = "txpInfo.bin: transcript offsets are not increasing inside the text"; return false; }
    }
    uint32_t bad = 0;
    for (int32_t v : SA) bad |= static_cast<uint32_t>(v < 0 || v >= n);
    if (bad) { err = "sa.bin: suffix array entry outside the text"; return false; }
  }
  txpLens.resize(txpOffsets.size());
  for (size_t i = 0; i + 1 < txpOffsets.size(); ++i) txpLens[i] = (txpOffsets[i + 1] - 1) - txpOffsets[i];
  txpLens.back() = (static_cast<int32_t>(SA.size()) - 1) - txpOffsets.back();

  if (!perfectHash) {
    File f(dir + "hash.bin");
    if (!f.ok()) { err = "cannot open hash.bin"; return false; }
    uint64_t fsz = f.size(), hdr = 0, magic = 0, tableSize = 0, numBuckets = 0;
    if (!read32or64(f, magic, hdr) || !read32or64(f, tableSize, hdr) || !read32or64(f, numBuckets, hdr)) { err = "hash.bin: truncated"; return false; }
    if (magic != 0x24687531ULL) { err = "hash.bin: not a sparsepp table (bad magic)"; return false; }
    const uint64_t rec = bigSA ? 24 : sizeof(KmerRecord);  // {u64 k-mer, IndexT begin, IndexT end}
    if (numBuckets > fsz / rec || fsz < hdr + numBuckets * rec) { err = "hash.bin: truncated"; return false; }
    // Group occupancy bitmaps (one word per 32 or 64 buckets) sit between header and records; their size is
    // whatever is left.  Only the records matter for a re-laid-out device table.
    uint64_t meta = fsz - hdr - numBuckets * rec;
    if (meta != (tableSize + 31) / 32 * 4 && meta != (tableSize + 63) / 64 * 8) { err = "hash.bin: unexpected group-bitmap size"; return false; }
    std::fseek(f.f, static_cast<long>(hdr + meta), SEEK_SET);
    kmers.resize(numBuckets);
    if (!bigSA) {
      if (!f.read(kmers.data(), numBuckets * rec)) { err = "hash.bin: truncated"; return false; }
    } else {
      struct Wide { uint64_t kmer; int64_t begin, end; };
      std::vector<Wide> wide(std::min<uint64_t>(numBuckets, 1 << 20));
      for (uint64_t at = 0; at < numBuckets; at += wide.size()) {
        const uint64_t cnt = std::min<uint64_t>(wide.size(), numBuckets - at);
        if (!f.read(wide.data(), cnt * rec)) { err = "hash.bin: truncated"; return false; }
        for (uint64_t i = 0; i < cnt; ++i) {
          if (wide[i].begin < 0 || wide[i].end < wide[i].begin || wide[i].end > static_cast<int64_t>(SA.size())) { err = "hash.bin: k-mer interval outside the suffix array"; return false; }
          kmers[at + i] = KmerRecord{wide[i].kmer, static_cast<int32_t>(wide[i].begin), static_cast<int32_t>(wide[i].end)};
        }
      }
    }
    {
      const int64_t n = static_cast<int64_t>(SA.size());
      uint32_t bad = 0;
      for (const auto& r : kmers) bad |= static_cast<uint32_t>(r.begin < 0 || r.end < r.begin || r.end > n);
      if (bad) { err = "hash.bin: k-mer interval outside the suffix array"; return false; }
    }
  } else {
    // ---- hash_info.bph: boomphf::mphf::save (include/BooPHF.hpp:1172-1197)
    File f(dir + "hash_info.bph");
    if (!f.ok()) { err = "cannot open hash_info.bph"; return false; }
    if (!f.get(phf.gamma) || !f.get(phf.nbLevels) || !f.get(phf.lastBitsetRank) || !f.get(phf.nelem)) { err = "hash_info.bph: truncated"; return false; }
    if (phf.nbLevels < 2 || phf.nbLevels > 64) { err = "hash_info.bph: implausible level count"; return false; }
    phf.levels.resize(static_cast<size_t>(phf.nbLevels));
    for (auto& lv : phf.levels) {
      uint64_t nchar = 0, nranks = 0;
      const uint64_t bsz = f.size();
      if (!f.get(lv.sizeBits) || !f.get(nchar)) { err = "hash_info.bph: truncated"; return false; }
      if (nchar > bsz / 8) { err = "hash_info.bph: bitset size exceeds the file size"; return false; }
      lv.bits.resize(nchar);
      if (!f.read(lv.bits.data(), nchar * 8) || !f.get(nranks)) { err = "hash_info.bph: truncated"; return false; }
      if (nranks > bsz / 8) { err = "hash_info.bph: rank table size exceeds the file size"; return false; }
      lv.ranks.resize(nranks);
      if (!f.read(lv.ranks.data(), nranks * 8)) { err = "hash_info.bph: truncated"; return false; }
    }
    {  // level domains are not stored: mphf::load recomputes them with pow() (:1219-1230); same libm here
      const double nelemD = static_cast<double>(phf.nelem);
      const double proba = 1.0 - std::pow(((phf.gamma * nelemD - 1) / (phf.gamma * nelemD)), static_cast<double>(phf.nelem - 1));
      const uint64_t hashDomain = static_cast<uint64_t>(std::ceil(nelemD * phf.gamma));
      for (int ii = 0; ii < phf.nbLevels; ++ii) {
        uint64_t d = ((static_cast<uint64_t>(static_cast<double>(hashDomain) * std::pow(proba, ii)) + 63) / 64) * 64;
        if (d == 0) d = 64;
        phf.levels[static_cast<size_t>(ii)].hashDomain = d;
      }
    }
    uint64_t nfinal = 0;
    if (!f.get(nfinal)) { err = "hash_info.bph: truncated"; return false; }
    if (nfinal > f.size() / 16) { err = "hash_info.bph: final-hash size exceeds the file size"; return false; }
    phf.finalHash.resize(nfinal);
    for (auto& kv : phf.finalHash)
      if (!f.get(kv.first) || !f.get(kv.second)) { err = "hash_info.bph: truncated"; return false; }
    std::sort(phf.finalHash.begin(), phf.finalHash.end());
    // ---- hash_info.val: FrugalBooMap::save (include/FrugalBooMap.hpp:199-213)
    File v(dir + "hash_info.val");
    if (!v.ok()) { err = "cannot open hash_info.val"; return false; }
    uint64_t n = 0;
    if (!v.get(n)) { err = "hash_info.val: truncated"; return false; }
    if (n > v.size() / (bigSA ? 8 : 4)) { err = "hash_info.val: element count exceeds the file size"; return false; }
    phf.data.resize(n);
    if (!bigSA) {
      if (!v.read(phf.data.data(), n * 4)) { err = "hash_info.val: truncated"; return false; }
    } else {
      std::vector<int64_t> wide(n);
      if (!v.read(wide.data(), n * 8)) { err = "hash_info.val: truncated"; return false; }
      for (uint64_t i = 0; i < n; ++i) {
        if (wide[i] < 0 || wide[i] >= static_cast<int64_t>(SA.size())) { err = "hash_info.val: interval start outside the suffix array"; return false; }
        phf.data[i] = static_cast<int32_t>(wide[i]);
      }
    }
    if (!v.get(n)) { err = "hash_info.val: truncated"; return false; }
    if (n > v.size()) { err = "hash_info.val: element count exceeds the file size"; return false; }
    phf.lens.resize(n);
    if (!v.read(phf.lens.data(), n)) { err = "hash_info.val: truncated"; return false; }
    if (phf.lens.size() != phf.data.size()) { err = "hash_info.val: data_/lens_ size mismatch"; return false; }
    uint64_t hdr = 0, magic = 0, tableSize = 0, numBuckets = 0;
    if (!read32or64(v, magic, hdr) || !read32or64(v, tableSize, hdr) || !read32or64(v, numBuckets, hdr) || magic != 0x24687531ULL) {
      err = "hash_info.val: bad overflow table";
      return false;
    }
    long here = std::ftell(v.f);
    uint64_t remaining = v.size() - static_cast<uint64_t>(here);
    const uint64_t orec = bigSA ? 16 : 8;  // {IndexT start, IndexT length}
    if (numBuckets > remaining / orec) { err = "hash_info.val: truncated overflow table"; return false; }
    std::fseek(v.f, static_cast<long>(static_cast<uint64_t>(here) + (remaining - numBuckets * orec)), SEEK_SET);
    phf.overflow.resize(numBuckets);
    for (auto& kv : phf.overflow) {
      if (!bigSA) {
        if (!v.get(kv.first) || !v.get(kv.second)) { err = "hash_info.val: truncated"; return false; }
      } else {
        int64_t a = 0, b = 0;
        if (!v.get(a) || !v.get(b)) { err = "hash_info.val: truncated"; return false; }
        if (a < 0 || b < 0 || a >= (1ll << 31) || b >= (1ll << 31)) return tooBig();
        kv.first = static_cast<int32_t>(a); kv.second = static_cast<int32_t>(b);
      }
    }
    std::sort(phf.overflow.begin(), phf.overflow.end());
    {
      const int64_t n = static_cast<int64_t>(SA.size());
      uint32_t bad = 0;
      for (int32_t v2 : phf.data) bad |= static_cast<uint32_t>(v2 < 0 || v2 >= n);
      if (bad) { err = "hash_info.val: interval start outside the suffix array"; return false; }
    }
  }
  return true;
}

} // namespace rapmap_b200
