#include "index_loader.hpp"

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
  if (bigSA) {
    // 64-bit suffix arrays (text > 2^31) are row f4 of SURVEY.md §8; refuse loudly rather than truncate.
    err = "BigSA (64-bit suffix array) indexes are not supported by the device path yet";
    return false;
  }

  {
    File f(dir + "sa.bin");
    if (!f.ok()) { err = "cannot open sa.bin"; return false; }
    uint64_t n = 0;
    if (!f.get(n)) { err = "sa.bin: truncated"; return false; }
    SA.resize(n);
    if (!f.read(SA.data(), n * sizeof(int32_t))) { err = "sa.bin: truncated"; return false; }
  }
  {
    File f(dir + "txpInfo.bin");
    if (!f.ok()) { err = "cannot open txpInfo.bin"; return false; }
    uint64_t n = 0;
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    txpNames.resize(n);
    for (auto& s : txpNames) {
      uint64_t l = 0;
      if (!f.get(l)) { err = "txpInfo.bin: truncated"; return false; }
      s.resize(l);
      if (!f.read(&s[0], l)) { err = "txpInfo.bin: truncated"; return false; }
    }
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    txpOffsets.resize(n);
    if (!f.read(txpOffsets.data(), n * sizeof(int32_t))) { err = "txpInfo.bin: truncated"; return false; }
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    text.resize(n);
    if (!f.read(&text[0], n)) { err = "txpInfo.bin: truncated"; return false; }
    if (!f.get(n)) { err = "txpInfo.bin: truncated"; return false; }
    txpCompleteLens.resize(n);
    if (!f.read(txpCompleteLens.data(), n * sizeof(uint32_t))) { err = "txpInfo.bin: truncated"; return false; }
  }
  if (txpOffsets.empty() || txpOffsets.size() != txpNames.size()) { err = "txpInfo.bin: inconsistent transcript tables"; return false; }
  if (SA.size() != text.size()) { err = "sa.bin / txpInfo.bin: suffix array and text lengths differ"; return false; }
  {
    File f(dir + "rsd.bin");
    if (!f.ok()) { err = "cannot open rsd.bin"; return false; }
    if (!f.get(numBits)) { err = "rsd.bin: truncated"; return false; }
    uint64_t nbytes = (numBits + 7) / 8;
    rsdBits.assign((numBits + 63) / 64 + 1, 0);
    if (!f.read(rsdBits.data(), nbytes)) { err = "rsd.bin: truncated"; return false; }
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
    const uint64_t rec = sizeof(KmerRecord);
    if (fsz < hdr + numBuckets * rec) { err = "hash.bin: truncated"; return false; }
    // Group occupancy bitmaps (one word per 32 or 64 buckets) sit between header and records; their size is
    // whatever is left.  Only the records matter for a re-laid-out device table.
    uint64_t meta = fsz - hdr - numBuckets * rec;
    if (meta != (tableSize + 31) / 32 * 4 && meta != (tableSize + 63) / 64 * 8) { err = "hash.bin: unexpected group-bitmap size"; return false; }
    std::fseek(f.f, static_cast<long>(hdr + meta), SEEK_SET);
    kmers.resize(numBuckets);
    if (!f.read(kmers.data(), numBuckets * rec)) { err = "hash.bin: truncated"; return false; }
  } else {
    // -p indexes: FrugalBooMap::find (include/FrugalBooMap.hpp:149-167) can only ever return
    // "k-mer of the text -> its SA range" (it verifies the key against the text at SA[start]).  The same
    // (k-mer, [begin,end)) set is recovered here by one scan of SA + text, and served by the device hash
    // table.  hash_info.bph / hash_info.val are not read.  (On-device BooPHF probing: next round, DESIGN.md.)
    const int64_t n = static_cast<int64_t>(SA.size());
    const int kk = static_cast<int>(k);
    bool have = false;
    uint64_t prev = 0;
    int32_t start = 0;
    for (int64_t i = 0; i <= n; ++i) {
      uint64_t w = 0;
      bool valid = false;
      if (i < n && static_cast<int64_t>(SA[i]) + kk <= n) {
        valid = true;
        const char* s = text.data() + SA[i];
        for (int j = 0; j < kk; ++j) {
          int c = code(s[j]);
          if (c < 0) { valid = false; break; }
          w = (w << 2) | static_cast<uint64_t>(c);
        }
      }
      if (have && (!valid || w != prev)) { kmers.push_back({prev, start, static_cast<int32_t>(i)}); have = false; }
      if (valid && !have) { prev = w; start = static_cast<int32_t>(i); have = true; }
    }
  }
  return true;
}

} // namespace rapmap_b200
