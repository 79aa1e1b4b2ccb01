// CPU ORACLE — TEST INFRASTRUCTURE ONLY (see quasimap_oracle.hpp for scope and parity status).
#include "quasimap_oracle.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <tuple>

namespace oracle {

// =================================================================================================
// Index loading — on-disk formats of SURVEY.md §5.2 (writers: src/RapMapSAIndexer.cpp:109-110,
// :694-733, :791-818; sparsepp serialisation include/sparsepp/spp.h:2355-2366,1603-1613,
// include/SparseHashSerializer.hpp:29-47)
// =================================================================================================
namespace {
struct Reader {
  std::ifstream f;
  explicit Reader(const std::string& p) : f(p, std::ios::binary) {}
  bool ok() const { return static_cast<bool>(f); }
  template <class T> T get() { T v{}; f.read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }
  void raw(void* p, size_t n) { f.read(static_cast<char*>(p), static_cast<std::streamsize>(n)); }
};
bool jsonBool(const std::string& txt, const std::string& key, bool def) {
  auto p = txt.find("\"" + key + "\"");
  if (p == std::string::npos) return def;
  p = txt.find(':', p);
  auto t = txt.find("true", p), f = txt.find("false", p);
  auto e = txt.find_first_of(",}", p);
  if (t != std::string::npos && t < e) return true;
  if (f != std::string::npos && f < e) return false;
  return def;
}
long jsonInt(const std::string& txt, const std::string& key, long def) {
  auto p = txt.find("\"" + key + "\"");
  if (p == std::string::npos) return def;
  p = txt.find(':', p);
  return std::strtol(txt.c_str() + p + 1, nullptr, 10);
}
uint32_t be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
} // namespace

bool Index::load(const std::string& dirIn, std::string* err) {
  std::string dir = dirIn;
  if (!dir.empty() && dir.back() != '/') dir += '/';
  auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
  {
    std::ifstream h(dir + "header.json");
    if (!h) return fail("cannot open header.json");
    std::stringstream ss; ss << h.rdbuf();
    std::string t = ss.str();
    k = static_cast<int>(jsonInt(t, "KmerLen", 31));
    bigSA = jsonBool(t, "BigSA", false);
    perfectHash = jsonBool(t, "PerfectHash", false);
  }
  {
    Reader r(dir + "sa.bin");
    if (!r.ok()) return fail("cannot open sa.bin");
    uint64_t n = r.get<uint64_t>();
    SA.resize(n);
    if (bigSA) r.raw(SA.data(), n * 8);
    else { std::vector<int32_t> t(n); r.raw(t.data(), n * 4); for (uint64_t i = 0; i < n; ++i) SA[i] = t[i]; }
  }
  {
    Reader r(dir + "txpInfo.bin");
    if (!r.ok()) return fail("cannot open txpInfo.bin");
    uint64_t n = r.get<uint64_t>();
    txpNames.resize(n);
    for (auto& s : txpNames) { uint64_t l = r.get<uint64_t>(); s.resize(l); r.raw(&s[0], l); }
    n = r.get<uint64_t>();
    txpOffsets.resize(n);
    if (bigSA) r.raw(txpOffsets.data(), n * 8);
    else { std::vector<int32_t> t(n); r.raw(t.data(), n * 4); for (uint64_t i = 0; i < n; ++i) txpOffsets[i] = t[i]; }
    n = r.get<uint64_t>();
    seq.resize(n);
    r.raw(&seq[0], n);
    n = r.get<uint64_t>();
    txpCompleteLens.resize(n);
    r.raw(txpCompleteLens.data(), n * 4);
  }
  {
    Reader r(dir + "rsd.bin");
    if (!r.ok()) return fail("cannot open rsd.bin");
    uint64_t nbits = r.get<uint64_t>();
    uint64_t nbytes = (nbits + 7) / 8;
    bits.assign((nbits + 63) / 64 + 1, 0);
    r.raw(bits.data(), nbytes);
    rankBlock.assign(bits.size() + 1, 0);
    for (size_t i = 0; i < bits.size(); ++i) rankBlock[i + 1] = rankBlock[i] + __builtin_popcountll(bits[i]);
  }
  // txpLens: src/RapMapSAIndex.cpp:151-163
  txpLens.resize(txpOffsets.size());
  for (size_t i = 0; i + 1 < txpOffsets.size(); ++i) txpLens[i] = (txpOffsets[i + 1] - 1) - txpOffsets[i];
  if (!txpOffsets.empty()) txpLens.back() = (static_cast<int64_t>(SA.size()) - 1) - txpOffsets.back();

  if (!perfectHash) {
    std::ifstream f(dir + "hash.bin", std::ios::binary);
    if (!f) return fail("cannot open hash.bin");
    f.seekg(0, std::ios::end);
    uint64_t fsz = static_cast<uint64_t>(f.tellg());
    f.seekg(0);
    auto rd3264 = [&](uint64_t& v) {
      unsigned char b[8];
      f.read(reinterpret_cast<char*>(b), 4);
      uint32_t x = be32(b);
      if (x != 0xFFFFFFFFu) { v = x; return 4; }
      f.read(reinterpret_cast<char*>(b), 8);
      v = (uint64_t(be32(b)) << 32) | be32(b + 4);
      return 12;
    };
    uint64_t magic, tableSize, numBuckets;
    uint64_t hdr = 0;
    hdr += rd3264(magic); hdr += rd3264(tableSize); hdr += rd3264(numBuckets);
    if (magic != 0x24687531ULL) return fail("hash.bin: bad sparsepp magic");
    uint64_t recSize = 8 + (bigSA ? 16 : 8);
    uint64_t meta = fsz - hdr - numBuckets * recSize; // group bitmaps: skipped, only the records matter
    f.seekg(static_cast<std::streamoff>(hdr + meta));
    khash.reserve(numBuckets * 2);
    for (uint64_t i = 0; i < numBuckets; ++i) {
      uint64_t key; f.read(reinterpret_cast<char*>(&key), 8);
      int64_t b, e;
      if (bigSA) { f.read(reinterpret_cast<char*>(&b), 8); f.read(reinterpret_cast<char*>(&e), 8); }
      else { int32_t b4, e4; f.read(reinterpret_cast<char*>(&b4), 4); f.read(reinterpret_cast<char*>(&e4), 4); b = b4; e = e4; }
      khash.emplace(key, std::make_pair(b, e));
    }
    if (!f) return fail("hash.bin: truncated");
  } else {
    // Perfect-hash index (include/FrugalBooMap.hpp:149-167): find(key) = MPHF slot -> verify the k-mer at
    // SA[data_[slot]] equals the key -> interval [start, start+len).  The set of (key -> interval) pairs it can
    // return is exactly "every distinct valid k-mer of the text -> its SA range" (src/RapMapSAIndexer.cpp:128-215
    // inserts the same pairs buildHash :261-443 does), so the oracle rebuilds that map directly from SA + text.
    int64_t n = static_cast<int64_t>(SA.size());
    uint64_t prev = 0; bool havePrev = false; int64_t start = 0;
    for (int64_t i = 0; i <= n; ++i) {
      uint64_t w = 0; bool valid = false;
      if (i < n && SA[i] + k <= static_cast<int64_t>(seq.size())) valid = encodeKmer(seq.data() + SA[i], k, w);
      if (havePrev && (!valid || w != prev)) { khash.emplace(prev, std::make_pair(start, i)); havePrev = false; }
      if (valid && !havePrev) { prev = w; start = i; havePrev = true; }
    }
  }
  return true;
}

int64_t Index::transcriptAtPosition(int64_t p) const {
  uint64_t word = static_cast<uint64_t>(p) / 64;
  return static_cast<int64_t>(rankBlock[word] + __builtin_popcountll(bits[word] & ((1ULL << (p % 64)) - 1)));
}

// =================================================================================================
// k-mer helpers
// =================================================================================================
static inline int baseCode(char c) { // include/Kmer.hpp:40-51
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}
bool encodeKmer(const char* s, int k, uint64_t& w) {
  w = 0;
  int shift = 2 * k - 2;
  for (int i = 0; i < k; ++i, shift -= 2) {
    int c = baseCode(s[i]);
    if (c < 0) return false;
    w |= (static_cast<uint64_t>(c) << shift);
  }
  return true;
}
uint64_t kmerRC(uint64_t w, int k) {
  w = ((w >> 2) & 0x3333333333333333ULL) | ((w & 0x3333333333333333ULL) << 2);
  w = ((w >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((w & 0x0F0F0F0F0F0F0F0FULL) << 4);
  w = ((w >> 8) & 0x00FF00FF00FF00FFULL) | ((w & 0x00FF00FF00FF00FFULL) << 8);
  w = ((w >> 16) & 0x0000FFFF0000FFFFULL) | ((w & 0x0000FFFF0000FFFFULL) << 16);
  w = (w >> 32) | (w << 32);
  return (~w) >> (2 * (32 - k));
}
static inline bool isHomopolymer(uint64_t w, int k) { // include/Kmer.hpp:484-487
  uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
  uint64_t nuc = w & 0x3;
  return w == (mask & ((w << 2) | nuc));
}
void reverseRead(const std::string& seq, std::string& out) { // src/RapMapUtils.cpp:107-128, table :63-72
  size_t n = seq.size();
  out.resize(n);
  for (size_t i = 0; i < n; ++i) {
    char c = seq[n - 1 - i], r;
    switch (c) {
      case 'A': case 'a': r = 'T'; break;
      case 'C': case 'c': r = 'G'; break;
      case 'G': case 'g': r = 'C'; break;
      case 'T': case 't': case 'U': case 'u': r = 'A'; break;
      default: r = 'N';
    }
    out[i] = r;
  }
}

QuasiAlignment::QuasiAlignment() : chainScore(std::numeric_limits<double>::lowest()) {}

// =================================================================================================
// Mapper
// =================================================================================================
Mapper::Mapper(const Index& i, const Opts& opts) : idx(i), o(opts) {
  // src/RapMapSAMapper.cpp:385-417
  disableNIP_ = o.sensitive;
  strictCheck_ = o.strictCheck;
  covReq_ = (o.quasiCov > 0.0) ? o.quasiCov : 0.0;
  doChaining_ = o.selAln;
  considerMultiPos_ = o.selAln;
  if (doChaining_) {
    consensusFraction_ = static_cast<float>((o.consensusSlack == 0.0) ? 1.0 : (1.0 - o.consensusSlack));
    strictCheckSlack_ = 1;                                   // include/SACollector.hpp:64
    if (o.maxMMPExtension > 0) maxMMPExtension_ = o.maxMMPExtension;  // :71
  }
}

struct Mapper::Impl {
  // ---------------------------------------------------------------------------------------------
  // include/SASearcher.hpp:87-309
  // ---------------------------------------------------------------------------------------------
  static std::tuple<int64_t, int64_t, int64_t> extendSearchNaive(Mapper& M, int64_t lbIn, int64_t ubIn, int64_t startAt,
                                                                 const char* qb, const char* qe) {
    const auto& SA = M.idx.SA;
    const std::string& seq = M.idx.seq;
    int64_t m = qe - qb;
    int64_t n = static_cast<int64_t>(seq.size());
    const char* sb = seq.data();
    if (ubIn - lbIn == 2) { // :109-126
      lbIn += 1;
      int64_t i = startAt;
      ++M.ops.saProbes;
      while (i < m && SA[lbIn] + i < n) {
        ++M.ops.textCmp;
        char qc = static_cast<char>(::toupper(qb[i]));
        if (qc != sb[SA[lbIn] + i]) break;
        ++i;
      }
      return std::make_tuple(lbIn, ubIn, i);
    }
    int64_t l = lbIn, r = ubIn;
    int64_t lcpLP = startAt, lcpRP = startAt;
    int64_t c = 0, i = 0;
    int64_t prevILow = startAt, prevIHigh = startAt;
    int64_t maxLen = startAt;
    bool plt = true;
    while (true) { // :150-209
      c = (l + r) / 2;
      ++M.ops.saProbes;
      plt = true;
      i = std::min(lcpLP, lcpRP);
      while (i < m && SA[c] + i < n) {
        ++M.ops.textCmp;
        char qc = static_cast<char>(::toupper(qb[i]));
        char tc = sb[SA[c] + i];
        if (qc < tc) {
          if (i > prevIHigh) prevIHigh = i;
          break;
        } else if (qc > tc) {
          if (i > prevILow) prevILow = i;
          plt = false;
          break;
        }
        ++i;
      }
      if (i == m || SA[c] + i == n) {
        if (i > prevIHigh) prevIHigh = i;
      }
      if (plt) {
        if (c == l + 1) { maxLen = std::max(std::max(i, prevILow), prevIHigh); break; }
        r = c; lcpRP = i;
      } else {
        if (c == r - 1) { maxLen = std::max(std::max(i, prevILow), prevIHigh); break; }
        l = c; lcpLP = i;
      }
    }
    m = maxLen + 1; // :212
    auto boundSearch = [&](int64_t lo, int64_t hi, char sentinel) -> int64_t { // :224-258 / :270-304
      int64_t l = lo, r = hi, lcpLP = startAt, lcpRP = startAt, c = 0, i = startAt;
      while (true) {
        c = (l + r) / 2;
        ++M.ops.saProbes;
        bool plt = true;
        i = std::min(lcpLP, lcpRP);
        while (i < m && SA[c] + i < n) {
          ++M.ops.textCmp;
          char qc = (i < m - 1) ? static_cast<char>(::toupper(qb[i])) : sentinel;
          char tc = sb[SA[c] + i];
          if (qc < tc) break;
          else if (qc > tc) { plt = false; break; }
          ++i;
        }
        if (plt) {
          if (c == l + 1) return c;
          r = c; lcpRP = i;
        } else {
          if (c == r - 1) return r;
          l = c; lcpLP = i;
        }
      }
    };
    int64_t b1 = boundSearch(lbIn, ubIn, '#');
    int64_t b2 = boundSearch(b1 - 1, ubIn, '{');
    if (b1 == b2) b2 += 1; // :307
    return std::make_tuple(b1, b2, maxLen);
  }

  // include/SASearcher.hpp:318-334
  static int64_t lce(Mapper& M, int64_t p1, int64_t p2, int64_t startAt, int64_t stopAt) {
    const std::string& seq = M.idx.seq;
    const auto& SA = M.idx.SA;
    int64_t textLen = static_cast<int64_t>(seq.size());
    int64_t len = startAt;
    int64_t o1 = SA[p1] + startAt, o2 = SA[p2] + startAt;
    int64_t maxIndex = std::max(o1, o2);
    while (maxIndex + len < textLen && seq[o1 + len] == seq[o2 + len]) {
      if (seq[o1 + len] == '$') break;
      if (len >= stopAt) break;
      ++len;
    }
    return len;
  }

  enum HitStatus { ABSENT = -1, UNTESTED = 0, PRESENT = 1 };
  struct KmerDirScore { uint64_t kmer; int32_t kpos; int fwdScore, rcScore; };
  using KmerScoreVec = std::vector<KmerDirScore>;

  static const std::pair<int64_t, int64_t>* find(Mapper& M, uint64_t w) { ++M.ops.hashFind; return M.idx.find(w); }

  // include/SACollector.hpp:366-431.  merKnown / compKnown: lookup result already available.
  static void spotCheck(Mapper& M, uint64_t mer, size_t pos, size_t readLen, bool merKnown, bool merPresent,
                        bool isRC, uint32_t& strandHits, uint32_t& otherStrandHits, KmerScoreVec& kmerScores) {
    int k = M.idx.k;
    uint64_t comp = kmerRC(mer, k);
    bool present = merKnown ? merPresent : (find(M, mer) != nullptr);
    bool compPresent = find(M, comp) != nullptr;
    int status = present ? PRESENT : ABSENT;
    int compStatus = compPresent ? PRESENT : ABSENT;
    if (present) ++strandHits;
    if (compPresent) ++otherStrandHits;
    int fwdStatus = isRC ? compStatus : status;
    int rcStatus = isRC ? status : compStatus;
    if (M.strictCheck_) {
      if (isRC) { pos = readLen - pos - k; mer = comp; }
      kmerScores.push_back({mer, static_cast<int32_t>(pos), fwdStatus, rcStatus});
    }
  }

  // include/SACollector.hpp:441-677
  static void getSAHits(Mapper& M, const std::string& read, size_t startPos, const std::pair<int64_t, int64_t>* startInterval,
                        size_t& cov, uint32_t& strandHits, uint32_t& otherStrandHits, std::vector<SAIntervalHit>& saInts,
                        KmerScoreVec& kmerScores, bool isRC) {
    const int k = M.idx.k;
    const int64_t readLen = static_cast<int64_t>(read.size());
    const char* rs = read.data();
    int64_t rb = 0;
    int64_t lb = 0, ub = 0;
    int64_t matchedLen = 0;
    bool lastSearch = false;
    size_t prevMMPEnd = 0;
    uint64_t mer = 0;
    bool skipSetup = (startInterval != nullptr);
    if (skipSetup) { rb = static_cast<int64_t>(startPos); lb = startInterval->first; ub = startInterval->second; }
    while (skipSetup || rb + k <= readLen) {
      bool haveHit = false;
      if (skipSetup) {
        haveHit = true;
        skipSetup = false;
      } else {
        int64_t pos = rb;
        bool validMer = encodeKmer(rs + pos, k, mer);
        if (!validMer) { // :505-516
          size_t invalidPos = read.find_first_of("Nn", static_cast<size_t>(pos));
          if (invalidPos < static_cast<size_t>(pos + k)) { rb = static_cast<int64_t>(invalidPos) + 1; continue; }
        }
        if (isHomopolymer(mer, k)) { rb += 1; continue; } // :520-536
        auto* it = find(M, mer);
        if (it) {
          spotCheck(M, mer, static_cast<size_t>(pos), static_cast<size_t>(readLen), true, true, isRC, strandHits, otherStrandHits, kmerScores);
          lb = it->first; ub = it->second;
          haveHit = true;
        } else {
          spotCheck(M, mer, static_cast<size_t>(pos), static_cast<size_t>(readLen), true, false, isRC, strandHits, otherStrandHits, kmerScores);
          rb += 1; // :673
          continue;
        }
      }
      if (haveHit) {
        lb = std::max<int64_t>(0, lb - 1); // :553
        bool firstAttempt = M.doChaining_ ? (rb == 0) : true;
        int64_t endPos = firstAttempt ? readLen : std::min<int64_t>(rb + k + M.maxMMPExtension_, readLen);
        int64_t lbP = lb, ubP = ub;
        std::tie(lb, ub, matchedLen) = extendSearchNaive(M, lb, ub, k, rs + rb, rs + endPos);
        if (M.doChaining_ && firstAttempt && !(matchedLen >= readLen) && matchedLen >= static_cast<int64_t>(k + M.maxMMPExtension_)) { // :568-575
          firstAttempt = false;
          lb = lbP; ub = ubP;
          endPos = std::min<int64_t>(rb + k + M.maxMMPExtension_, readLen);
          std::tie(lb, ub, matchedLen) = extendSearchNaive(M, lb, ub, k, rs + rb, rs + endPos);
        }
        int64_t diff = ub - lb;
        if (ub > lb && diff < M.maxInterval_) { // :578
          saInts.push_back({lb, ub, static_cast<uint32_t>(matchedLen), static_cast<uint32_t>(rb), isRC});
          ++M.ops.intervals;
          size_t matchOffset = static_cast<size_t>(rb);
          size_t correction = 0;
          if (prevMMPEnd > matchOffset) correction = prevMMPEnd - matchOffset;
          cov += (static_cast<size_t>(matchedLen) - correction);
          prevMMPEnd = matchOffset + static_cast<size_t>(matchedLen);
          if (rb + matchedLen < readLen) { // :599-616
            int64_t kmerPos = rb + matchedLen - (k - 1);
            uint64_t mm;
            if (encodeKmer(rs + kmerPos, k, mm)) {
              mer = mm;
              spotCheck(M, mm, static_cast<size_t>(kmerPos), static_cast<size_t>(readLen), false, false, isRC, strandHits, otherStrandHits, kmerScores);
            } else {
              mer = mm; // the reference decodes into the same `mer` object; value unused afterwards
            }
          }
        }
        if (lastSearch) return;                 // :623
        int64_t mismatchPos = rb + matchedLen;
        if (mismatchPos >= readLen) return;     // :630
        int64_t remaining = readLen - mismatchPos;
        int64_t lceLen = M.disableNIP_ ? matchedLen : lce(M, lb, ub - 1, matchedLen, remaining);
        int64_t skipMatch = mismatchPos - (k - 1);
        int64_t skipLCE = rb + lceLen - (k - 1);
        rb = std::max(skipMatch, skipLCE);
        if (!M.disableNIP_ && lceLen > matchedLen) {
          if (readLen > k) rb = std::min<int64_t>(readLen - k, rb);
        }
        if (rb + k == readLen) lastSearch = true; // :663
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // src/HitManager.cpp
  // ---------------------------------------------------------------------------------------------
  struct SATxpQueryPos { uint32_t pos, queryPos; bool queryRC; int32_t len; };
  struct ProcessedSAHit {
    std::vector<SATxpQueryPos> tqvec;
    bool active{false};
    uint32_t numActive{1};
    uint32_t lastActiveInterval{1};
  };
  using SAHitMap = std::map<int, ProcessedSAHit>;

  static float fastlog2(float x) { // :33-42
    union { float f; uint32_t i; } vx = {x};
    union { uint32_t i; float f; } mx = {(vx.i & 0x007FFFFF) | 0x3f000000};
    float y = vx.i;
    y *= 1.1920928955078125e-7f;
    return y - 124.22551499f - 1.498030302f * mx.f - 1.72587999f / (0.3520887068f + mx.f);
  }

  // :449-493
  static void intersectSAIntervalWithOutput(Mapper& M, const SAIntervalHit& h, uint32_t intervalCounter, int32_t maxSlack, SAHitMap& outHits) {
    const auto& SA = M.idx.SA;
    bool nonStrict = maxSlack > 0;
    for (int64_t i = h.begin; i != h.end; ++i) {
      ++M.ops.rankCalls;
      int64_t txpID = M.idx.transcriptAtPosition(SA[i]);
      auto it = outHits.find(static_cast<int>(txpID));
      bool inOutputSet = (it != outHits.end());
      int32_t txpOccCount = inOutputSet ? static_cast<int32_t>(it->second.numActive) : 0;
      int32_t slack = (static_cast<int32_t>(intervalCounter) - 1) - txpOccCount;
      if (nonStrict || slack <= maxSlack) {
        int64_t localPos = SA[i] - M.idx.txpOffsets[txpID];
        if (inOutputSet) {
          it->second.numActive += (it->second.lastActiveInterval == intervalCounter) ? 0 : 1;
          it->second.lastActiveInterval = intervalCounter;
          it->second.tqvec.push_back({static_cast<uint32_t>(localPos), h.queryPos, h.queryRC, static_cast<int32_t>(h.len)});
        } else {
          auto& oh = outHits[static_cast<int>(txpID)];
          oh.tqvec.push_back({static_cast<uint32_t>(localPos), h.queryPos, h.queryRC, static_cast<int32_t>(h.len)});
          oh.lastActiveInterval = intervalCounter;
        }
      }
    }
  }

  // :587-689 (strictFilter == mc.consistentHits == false always, src/RapMapSAMapper.cpp:409)
  static SAHitMap intersectSAHits(Mapper& M, std::vector<SAIntervalHit>& inHits, float consensusFraction) {
    SAHitMap outHits;
    if (inHits.size() < 2) return outHits;
    const auto& SA = M.idx.SA;
    int32_t sIn = static_cast<int32_t>(inHits.size());
    float requiredFrac = sIn * consensusFraction;
    int32_t requiredNumHits = sIn;
    int32_t maxSlack = 0;
    if (consensusFraction < 1.0) {
      requiredNumHits = std::max<int32_t>(1, static_cast<int32_t>(std::floor(requiredFrac)));
      maxSlack = sIn - requiredNumHits;
    }
    SAIntervalHit* minHit = &inHits[0];
    for (auto& h : inHits) if (h.span() < minHit->span()) minHit = &h;
    for (int64_t i = minHit->begin; i < minHit->end; ++i) {
      ++M.ops.rankCalls;
      int64_t g = SA[i];
      int64_t tid = M.idx.transcriptAtPosition(g);
      int64_t txpPos = g - M.idx.txpOffsets[tid];
      auto& oh = outHits[static_cast<int>(tid)];
      oh.tqvec.push_back({static_cast<uint32_t>(txpPos), minHit->queryPos, minHit->queryRC, static_cast<int32_t>(minHit->len)});
      oh.lastActiveInterval = 1;
    }
    uint32_t intervalCounter = 2;
    for (auto& h : inHits) {
      if (&h != minHit) { intersectSAIntervalWithOutput(M, h, intervalCounter, maxSlack, outHits); ++intervalCounter; }
    }
    size_t numActive = 0;
    for (auto& kv : outHits) {
      bool enough = (static_cast<int32_t>(kv.second.numActive) >= requiredNumHits);
      kv.second.active = enough;
      numActive += enough ? 1 : 0;
    }
    if (maxSlack > 0 && numActive == 0) for (auto& kv : outHits) kv.second.active = true;
    return outHits;
  }

  // :84-326
  static void collectHitsSimpleSA(Mapper& M, SAHitMap& processedHits, uint32_t readLen, int32_t maxDist,
                                  std::vector<QuasiAlignment>& hits, uint8_t mateStatus) {
    bool findBestChain = M.doChaining_;
    bool considerMultiPos = M.considerMultiPos_;
    std::vector<double> f;
    std::vector<int32_t> p;
    std::vector<int32_t> bestChainEndInds;
    for (auto& ph : processedHits) {
      if (!ph.second.active) continue;
      uint32_t tid = static_cast<uint32_t>(ph.first);
      if (findBestChain) {
        auto& hv = ph.second.tqvec;
        std::sort(hv.begin(), hv.end(), [](const SATxpQueryPos& p1, const SATxpQueryPos& p2) {
          auto r1 = p1.pos + p1.len, r2 = p2.pos + p2.len;
          auto q1 = p1.queryPos + p1.len, q2 = p2.queryPos + p2.len;
          return (r1 < r2) ? true : ((r2 < r1) ? false : (q1 < q2));
        });
        auto alpha = [](int32_t qdiff, int32_t rdiff, int32_t ilen) -> double {
          double score = ilen;
          double mindiff = (qdiff < rdiff) ? qdiff : rdiff;
          return (score < mindiff) ? score : mindiff;
        };
        auto beta = [maxDist](int32_t qdiff, int32_t rdiff, double avgseed) -> double {
          if (qdiff < 0 || (std::max(qdiff, rdiff) > maxDist)) return std::numeric_limits<double>::infinity();
          double l = qdiff - rdiff;
          int32_t al = std::abs(l);
          return (l == 0) ? 0.0 : (0.01 * avgseed * al + 0.5 * fastlog2(static_cast<float>(al)));
        };
        double bestScore = std::numeric_limits<double>::lowest();
        int32_t bestChainEnd = -1;
        double avgseed = 31.0;
        bestChainEndInds.clear(); f.clear(); p.clear();
        int32_t lastHitId = static_cast<int32_t>(hv.size()) - 1;
        for (int32_t i = 0; i < static_cast<int32_t>(hv.size()); ++i) {
          auto& hi = hv[i];
          auto qposi = hi.queryPos + hi.len;
          auto rposi = hi.pos + hi.len;
          p.push_back(i);
          f.push_back(static_cast<double>(hi.len));
          int32_t numRounds = 2;
          for (int32_t j = i - 1; j >= 0; --j) {
            auto& hj = hv[j];
            auto qposj = hj.queryPos + hj.len;
            auto rposj = hj.pos + hj.len;
            // NB: uint32 arithmetic converted to int32 at the lambda boundary, as in the reference
            int32_t qdiff = static_cast<int32_t>(qposi - qposj);
            int32_t rdiff = static_cast<int32_t>(rposi - rposj);
            double ext = f[j] + alpha(qdiff, rdiff, hi.len) - beta(qdiff, rdiff, avgseed);
            bool extendWithJ = (ext > f[i]);
            p[i] = extendWithJ ? j : p[i];
            f[i] = extendWithJ ? ext : f[i];
            if (p[i] < i) { numRounds--; if (numRounds <= 0) break; }
          }
          if (f[i] > bestScore) {
            bestScore = f[i]; bestChainEnd = i;
            if (considerMultiPos) { bestChainEndInds.clear(); bestChainEndInds.push_back(bestChainEnd); }
          } else if (considerMultiPos && f[i] == bestScore) {
            bestChainEndInds.push_back(i);
          }
        }
        if (!considerMultiPos) bestChainEndInds.push_back(bestChainEnd);
        size_t numDistinctOpt = 0;
        std::vector<int8_t> seen(f.size(), 0);
        std::vector<int32_t> startPositions; // indices into hv
        int32_t lastChainHit = bestChainEnd;
        for (int32_t bestChainEndInd : bestChainEndInds) {
          bool validChain = true;
          int32_t lastPtr = p[bestChainEndInd];
          while (lastPtr < bestChainEndInd) {
            if (seen[bestChainEndInd]) { validChain = false; break; }
            seen[bestChainEndInd] = 1;
            bestChainEndInd = lastPtr;
            lastPtr = p[bestChainEndInd];
          }
          if (seen[bestChainEndInd]) validChain = false;
          if (validChain) { ++numDistinctOpt; startPositions.push_back(lastPtr); }
        }
        {
          auto& s0 = hv[startPositions[0]];
          bool isFwd = !s0.queryRC;
          int32_t hitPos = static_cast<int32_t>(s0.pos - s0.queryPos);
          QuasiAlignment qa;
          qa.tid = tid; qa.pos = hitPos; qa.fwd = isFwd; qa.readLen = readLen;
          qa.mateIsFwd = true; qa.chainScore = bestScore; qa.mateStatus = mateStatus;
          qa.allPositions.push_back(hitPos);
          if (startPositions.size() > 1) {
            qa.hasMultiPos = true;
            for (size_t s = 1; s < startPositions.size(); ++s) {
              auto& sp = hv[startPositions[s]];
              qa.allPositions.push_back(static_cast<int32_t>(sp.pos - sp.queryPos));
            }
            std::sort(qa.allPositions.begin(), qa.allPositions.end());
          }
          hits.push_back(std::move(qa));
        }
        if (hv.size() > 1 && numDistinctOpt == 1 && lastChainHit == lastHitId) { // :283-306
          auto& lastHit = hv[lastHitId];
          int64_t queryRange = static_cast<int64_t>(lastHit.queryPos + lastHit.len) - hv[0].queryPos;
          int64_t refRange = static_cast<int64_t>(lastHit.pos + lastHit.len) - hv[0].pos;
          if (queryRange == refRange && queryRange == static_cast<int64_t>(readLen)) {
            if (mateStatus == SINGLE_END || mateStatus == PAIRED_END_LEFT) hits.back().chainLeft = UNGAPPED;
            else if (mateStatus == PAIRED_END_RIGHT) hits.back().chainRight = UNGAPPED;
          }
        }
      } else { // :308-322
        auto& tq = ph.second.tqvec;
        auto minIt = std::min_element(tq.begin(), tq.end(), [](const SATxpQueryPos& a, const SATxpQueryPos& b) { return a.pos < b.pos; });
        QuasiAlignment qa;
        qa.tid = tid; qa.pos = static_cast<int32_t>(minIt->pos - minIt->queryPos); qa.fwd = !minIt->queryRC;
        qa.readLen = readLen; qa.mateIsFwd = true; qa.mateStatus = mateStatus;
        qa.allPositions.push_back(qa.pos);
        hits.push_back(std::move(qa));
      }
    }
  }

  // :716-807
  static void collectFromSingleInterval(Mapper& M, std::vector<SAIntervalHit>& saInts, bool isFw, uint8_t mateStatus, uint32_t readLen,
                                        std::vector<QuasiAlignment>& outHits) {
    auto& h = saInts.front();
    size_t initialSize = outHits.size();
    const auto& SA = M.idx.SA;
    for (int64_t i = h.begin; i != h.end; ++i) {
      ++M.ops.rankCalls;
      int64_t g = SA[i];
      int64_t txpID = M.idx.transcriptAtPosition(g);
      int64_t pos = g - M.idx.txpOffsets[txpID];
      int32_t hitPos = static_cast<int32_t>(pos - h.queryPos);
      QuasiAlignment qa;
      qa.tid = static_cast<uint32_t>(txpID); qa.pos = hitPos; qa.fwd = isFw; qa.readLen = readLen;
      qa.mateIsFwd = true; qa.mateStatus = mateStatus; qa.allPositions.push_back(hitPos); qa.hasMultiPos = false;
      uint8_t cs = (h.len == readLen) ? PERFECT : REGULAR;
      if (mateStatus == PAIRED_END_LEFT || mateStatus == SINGLE_END) qa.chainLeft = cs;
      else if (mateStatus == PAIRED_END_RIGHT) qa.chainRight = cs;
      outHits.push_back(std::move(qa));
    }
    std::sort(outHits.begin() + initialSize, outHits.end(), [](const QuasiAlignment& a, const QuasiAlignment& b) {
      return (a.tid == b.tid) ? (a.pos < b.pos) : (a.tid < b.tid);
    });
    // mergeUnique (:769-793) or std::unique by tid (:799-803)
    size_t w = initialSize;
    for (size_t r = initialSize + 1; r < outHits.size(); ++r) {
      if (outHits[r].tid != outHits[w].tid) {
        ++w;
        if (w != r) outHits[w] = std::move(outHits[r]);
      } else if (M.considerMultiPos_) {
        outHits[w].hasMultiPos = true;
        outHits[w].allPositions.push_back(outHits[r].pos);
      }
    }
    if (outHits.size() > initialSize) outHits.resize(w + 1);
  }
};

// include/SACollector.hpp:108-362
bool Mapper::collect(const std::string& read, HitCollectorInfo& hc) {
  using I = Impl;
  const int k = idx.k;
  const size_t readLen = read.size();
  hc.readLen = readLen;
  hc.maxDist = static_cast<int32_t>(readLen);
  uint32_t fwdHit = 0, rcHit = 0;
  size_t fwdCov = 0, rcCov = 0;
  bool foundHit = false;
  bool useCoverageCheck = disableNIP_ && strictCheck_;
  I::KmerScoreVec kmerScores;
  const std::pair<int64_t, int64_t>* merIt = nullptr;
  const std::pair<int64_t, int64_t>* rcMerIt = nullptr;
  size_t rb = 0;
  size_t invalidPos = 0;
  size_t pos = 0;
  while (rb + k <= readLen) { // :167-237
    pos = rb;
    if (invalidPos != std::string::npos) {
      invalidPos = read.find_first_of("nN", pos);
      if (invalidPos <= pos + k) { rb = invalidPos + 1; continue; }
    }
    uint64_t mer;
    encodeKmer(read.data() + pos, k, mer);
    if (isHomopolymer(mer, k)) { rb += 1; continue; }
    uint64_t rcMer = kmerRC(mer, k);
    merIt = I::find(*this, mer);
    rcMerIt = I::find(*this, rcMer);
    if (merIt) {
      ++fwdHit;
      if (rcMerIt) {
        ++rcHit;
        if (strictCheck_) kmerScores.push_back({mer, static_cast<int32_t>(pos), I::PRESENT, I::PRESENT});
      } else if (strictCheck_) {
        kmerScores.push_back({mer, static_cast<int32_t>(pos), I::PRESENT, I::ABSENT});
      }
    }
    if (rcMerIt) {
      if (!fwdHit) {
        ++rcHit;
        if (strictCheck_) kmerScores.push_back({mer, static_cast<int32_t>(pos), I::ABSENT, I::PRESENT});
      }
    }
    if (fwdHit + rcHit > 0) { foundHit = true; break; }
    ++rb;
  }
  if (!foundHit) return false;

  bool didCheckFwd = false;
  if (fwdHit) {
    didCheckFwd = true;
    I::getSAHits(*this, read, rb, merIt, fwdCov, fwdHit, rcHit, hc.fwdSAInts, kmerScores, false);
  }
  bool checkRC = useCoverageCheck ? (rcHit > 0) : (rcHit >= fwdHit);
  if (checkRC) {
    std::string rcBuf;
    reverseRead(read, rcBuf);
    I::getSAHits(*this, rcBuf, 0, nullptr, rcCov, rcHit, fwdHit, hc.rcSAInts, kmerScores, true);
  }
  bool checkFwd = useCoverageCheck ? (fwdHit > 0) : (fwdHit >= rcHit);
  if (!didCheckFwd && checkFwd) {
    didCheckFwd = true;
    I::getSAHits(*this, read, 0, nullptr, fwdCov, fwdHit, rcHit, hc.fwdSAInts, kmerScores, false);
  }
  if (strictCheck_) { // :280-339
    if (useCoverageCheck) {
      if (fwdCov > rcCov + strictCheckSlack_) hc.rcSAInts.clear();
      else if (rcCov > fwdCov + strictCheckSlack_) hc.fwdSAInts.clear();
    } else {
      if (fwdHit > 0 && rcHit == 0) hc.rcSAInts.clear();
      else if (rcHit > 0 && fwdHit == 0) hc.fwdSAInts.clear();
      else {
        std::stable_sort(kmerScores.begin(), kmerScores.end(), [](const I::KmerDirScore& a, const I::KmerDirScore& b) { return a.kpos < b.kpos; });
        auto e = std::unique(kmerScores.begin(), kmerScores.end(), [](const I::KmerDirScore& a, const I::KmerDirScore& b) { return a.kpos == b.kpos; });
        int32_t fwdScore = 0, rcScore = 0;
        for (auto it = kmerScores.begin(); it != e; ++it) { fwdScore += it->fwdScore; rcScore += it->rcScore; }
        if (fwdScore > rcScore) hc.rcSAInts.clear();
        else if (rcScore > fwdScore) hc.fwdSAInts.clear();
      }
    }
  }
  if (covReq_ > 0.0 && disableNIP_) { // :343-358
    if (!hc.fwdSAInts.empty()) { double fr = fwdCov / static_cast<double>(readLen); if (fr < covReq_) hc.fwdSAInts.clear(); }
    if (!hc.rcSAInts.empty()) { double fr = rcCov / static_cast<double>(readLen); if (fr < covReq_) hc.rcSAInts.clear(); }
  }
  return foundHit;
}

// src/HitManager.cpp:691-882
void Mapper::hitsToMappingsSimple(uint8_t mateStatus, HitCollectorInfo& hc, std::vector<QuasiAlignment>& hits) {
  using I = Impl;
  uint32_t readLen = static_cast<uint32_t>(hc.readLen);
  size_t fwdStart = hits.size();
  if (hc.fwdSAInts.size() > 1) {
    auto ph = I::intersectSAHits(*this, hc.fwdSAInts, consensusFraction_);
    I::collectHitsSimpleSA(*this, ph, readLen, hc.maxDist, hits, mateStatus);
  } else if (hc.fwdSAInts.size() == 1) {
    I::collectFromSingleInterval(*this, hc.fwdSAInts, true, mateStatus, readLen, hits);
  }
  size_t fwdEnd = hits.size();
  size_t rcStart = fwdEnd;
  if (hc.rcSAInts.size() > 1) {
    auto ph = I::intersectSAHits(*this, hc.rcSAInts, consensusFraction_);
    I::collectHitsSimpleSA(*this, ph, readLen, hc.maxDist, hits, mateStatus);
  } else if (hc.rcSAInts.size() == 1) {
    I::collectFromSingleInterval(*this, hc.rcSAInts, false, mateStatus, readLen, hits);
  }
  size_t rcEnd = hits.size();
  if (fwdEnd > fwdStart && rcEnd > rcStart) { // :834-881
    std::inplace_merge(hits.begin() + fwdStart, hits.begin() + fwdEnd, hits.begin() + rcEnd,
                       [](const QuasiAlignment& a, const QuasiAlignment& b) {
                         return (a.tid == b.tid) ? a.chainScore > b.chainScore : a.tid < b.tid;
                       });
    size_t w = fwdStart;
    for (size_t r = fwdStart + 1; r < rcEnd; ++r) {
      if (hits[r].tid != hits[w].tid) {
        ++w;
        if (w != r) hits[w] = std::move(hits[r]);
      } else {
        hits[w].oppositeStrandPositions = hits[r].allPositions;
      }
    }
    hits.resize(w + 1);
  }
}

// =================================================================================================
// Mate merging — include/RapMapUtils.hpp:864-1264
// =================================================================================================
namespace {
enum class MergeResult : uint8_t { HAD_NONE, HAD_EMPTY_INTERSECTION, HAD_CONCORDANT, HAD_DISCORDANT, HAD_ONLY_LEFT, HAD_ONLY_RIGHT };

// :903-975 — (fwPos, rcPos, gap) of the closest rc-downstream-of-fwd pair, or false
bool findBestHitFWRC(const std::vector<int32_t>& fwdHits, const std::vector<int32_t>& rcHits, int32_t fwdReadLen,
                     int32_t& fwPos, int32_t& rcPos, int32_t& gapOut) {
  if (fwdHits.empty() || rcHits.empty()) return false;
  constexpr int32_t maxGap = std::numeric_limits<int32_t>::max();
  int32_t bestGap = maxGap;
  size_t bf = 0, br = 0;
  auto update = [&](size_t fi, size_t ri) {
    int32_t gap = (rcHits[ri] >= fwdHits[fi]) ? std::abs(rcHits[ri] - (fwdHits[fi] + fwdReadLen)) : maxGap;
    if (gap < bestGap) { bestGap = gap; bf = fi; br = ri; }
  };
  for (size_t fi = 0; fi < fwdHits.size(); ++fi) {
    int32_t p1 = fwdHits[fi];
    size_t lb = static_cast<size_t>(std::lower_bound(rcHits.begin(), rcHits.end(), p1) - rcHits.begin());
    if (lb == rcHits.size()) update(fi, lb - 1);
    else if (lb == 0) update(fi, lb);
    else { update(fi, lb); update(fi, lb - 1); }
  }
  if (bestGap == maxGap) return false;
  fwPos = fwdHits[bf]; rcPos = rcHits[br]; gapOut = bestGap;
  return true;
}

MergeResult mergeLeftRightHitsFuzzy(bool leftMatches, bool rightMatches, std::vector<QuasiAlignment>& leftHits,
                                    std::vector<QuasiAlignment>& rightHits, std::vector<QuasiAlignment>& jointHits,
                                    uint32_t maxNumHits, bool& tooManyHits, Counters& hctr) {
  MergeResult mergeRes = MergeResult::HAD_NONE;
  if (leftHits.empty()) {
    if (!leftMatches) {
      if (!rightHits.empty()) {
        jointHits.insert(jointHits.end(), rightHits.begin(), rightHits.end());
        hctr.seHits += rightHits.size();
        mergeRes = MergeResult::HAD_ONLY_RIGHT;
      }
    }
  } else if (rightHits.empty()) {
    if (!rightMatches) {
      jointHits.insert(jointHits.end(), leftHits.begin(), leftHits.end());
      hctr.seHits += leftHits.size();
      mergeRes = MergeResult::HAD_ONLY_LEFT;
    }
  } else {
    uint32_t sameTxpCount = 0;
    size_t li = 0, ri = 0, numHits = 0;
    while (li < leftHits.size() && ri < rightHits.size()) {
      uint32_t leftTxp = leftHits[li].tid, rightTxp = rightHits[ri].tid;
      if (leftTxp < rightTxp) { ++li; }
      else {
        if (!(rightTxp < leftTxp)) {
          ++sameTxpCount;
          auto& L = leftHits[li];
          auto& R = rightHits[ri];
          auto& leftFwdHits = L.fwd ? L.allPositions : L.oppositeStrandPositions;
          auto& leftRCHits = L.fwd ? L.oppositeStrandPositions : L.allPositions;
          auto& rightFwdHits = R.fwd ? R.allPositions : R.oppositeStrandPositions;
          auto& rightRCHits = R.fwd ? R.oppositeStrandPositions : R.allPositions;
          int32_t a1, a2, ag, b1, b2, bg;
          bool haveFWRC = findBestHitFWRC(leftFwdHits, rightRCHits, static_cast<int32_t>(L.readLen), a1, a2, ag);
          bool haveRCFW = findBestHitFWRC(rightFwdHits, leftRCHits, static_cast<int32_t>(R.readLen), b1, b2, bg);
          bool foundValidHit = false, leftFwd = false, rightFwd = false;
          int32_t bestGap = std::numeric_limits<int32_t>::max();
          int32_t leftPos = -1, rightPos = -1;
          if (haveFWRC) { leftPos = a1; rightPos = a2; bestGap = ag; leftFwd = true; rightFwd = false; foundValidHit = true; }
          if (haveRCFW) {
            if (bg < bestGap) { leftPos = b2; rightPos = b1; leftFwd = false; rightFwd = true; }
            foundValidHit = true;
          }
          if (foundValidHit) {
            int32_t startRead1 = std::max(leftPos, 0), startRead2 = std::max(rightPos, 0);
            bool read1First = startRead1 < startRead2;
            int32_t fragStartPos = read1First ? startRead1 : startRead2;
            int32_t fragEndPos = read1First ? (startRead2 + static_cast<int32_t>(R.readLen)) : (startRead1 + static_cast<int32_t>(L.readLen));
            uint32_t fragLen = static_cast<uint32_t>(fragEndPos - fragStartPos);
            QuasiAlignment qa;
            qa.tid = leftTxp; qa.pos = leftPos; qa.fwd = leftFwd; qa.readLen = L.readLen; qa.fragLen = fragLen; qa.isPaired = true;
            qa.mateLen = R.readLen; qa.matePos = rightPos; qa.mateIsFwd = rightFwd; qa.mateStatus = PAIRED_END_PAIRED;
            qa.chainLeft = L.chainLeft; qa.chainRight = R.chainRight;
            jointHits.push_back(std::move(qa));
            ++numHits;
            mergeRes = MergeResult::HAD_CONCORDANT;
            if (numHits > maxNumHits) { tooManyHits = true; break; }
          }
          ++li;
        }
        ++ri;
      }
    }
    if (tooManyHits) { jointHits.clear(); ++hctr.tooManyHits; }
    if (mergeRes == MergeResult::HAD_NONE) mergeRes = (sameTxpCount > 0) ? MergeResult::HAD_DISCORDANT : MergeResult::HAD_EMPTY_INTERSECTION;
  }
  if (!jointHits.empty()) hctr.peHits += jointHits.size();
  return mergeRes;
}

// :1185-1264
void mergeLeftRightHits(std::vector<QuasiAlignment>& leftHits, std::vector<QuasiAlignment>& rightHits,
                        std::vector<QuasiAlignment>& jointHits, uint32_t maxNumHits, bool& tooManyHits, Counters& hctr) {
  if (!leftHits.empty()) {
    if (!rightHits.empty()) {
      size_t li = 0, ri = 0, numHits = 0;
      while (li < leftHits.size() && ri < rightHits.size()) {
        uint32_t leftTxp = leftHits[li].tid, rightTxp = rightHits[ri].tid;
        if (leftTxp < rightTxp) { ++li; }
        else {
          if (!(rightTxp < leftTxp)) {
            auto& L = leftHits[li];
            auto& R = rightHits[ri];
            int32_t startRead1 = std::max(L.pos, 0), startRead2 = std::max(R.pos, 0);
            bool read1First = startRead1 < startRead2;
            int32_t fragStartPos = read1First ? startRead1 : startRead2;
            int32_t fragEndPos = read1First ? (startRead2 + static_cast<int32_t>(R.readLen)) : (startRead1 + static_cast<int32_t>(L.readLen));
            uint32_t fragLen = static_cast<uint32_t>(fragEndPos - fragStartPos);
            QuasiAlignment qa;
            qa.tid = leftTxp; qa.pos = startRead1; qa.fwd = L.fwd; qa.readLen = L.readLen; qa.fragLen = fragLen; qa.isPaired = true;
            qa.mateLen = R.readLen; qa.matePos = startRead2; qa.mateIsFwd = R.fwd; qa.mateStatus = PAIRED_END_PAIRED;
            qa.chainLeft = L.chainLeft; qa.chainRight = R.chainRight;
            jointHits.push_back(std::move(qa));
            ++numHits;
            if (numHits > maxNumHits) { tooManyHits = true; break; }
            ++li;
          }
          ++ri;
        }
      }
    }
    if (tooManyHits) { jointHits.clear(); ++hctr.tooManyHits; }
  }
  if (!jointHits.empty()) {
    hctr.peHits += jointHits.size();
  } else if (leftHits.size() + rightHits.size() > 0 && !tooManyHits) {
    hctr.seHits += leftHits.size() + rightHits.size();
    jointHits.insert(jointHits.end(), leftHits.begin(), leftHits.end());
    jointHits.insert(jointHits.end(), rightHits.begin(), rightHits.end());
  }
}
} // namespace

// =================================================================================================
// Selective alignment — include/SelectiveAlignmentUtils.hpp:260-373, src/ksw2pp/*
// =================================================================================================
namespace {
constexpr int32_t KSW_NEG_INF = -0x40000000;
inline uint8_t nt4(char c) { // src/ksw2pp/KSW2Aligner.cpp:61-72
  switch (c) {
    case 'A': case 'a': case 0: return 0;
    case 'C': case 'c': case 1: return 1;
    case 'G': case 'g': case 2: return 2;
    case 'T': case 't': case 3: return 3;
    default: return 4;
  }
}
} // namespace

// Lane-for-lane restatement of ksw_extz2_sse41 (src/ksw2pp/ksw2_extz2_sse.c:18-304) for
// flag = KSW_EZ_SCORE_ONLY, zdrop = -1, m = 5, end_bonus unused in score-only mode.
// The u/v/x/y/s/sf/qr byte arrays live in ONE zero-filled buffer with the reference's layout
// (:84-86) because the SSE code reads 16-lane blocks past tlen/qlen into its neighbours.
int32_t Mapper::kswExtzScore(const char* qs, int qlen, const char* ts, int tlen) {
  ++ops.kswCalls;
  const int m = 5;
  int8_t mat[25];
  { // src/ksw2pp/KSW2Aligner.cpp:74-96
    int a = o.matchScore, b = o.mismatchPenalty;
    a = a < 0 ? -a : a;
    b = b > 0 ? -b : b;
    for (int i = 0; i < m - 1; ++i) {
      for (int j = 0; j < m - 1; ++j) mat[i * m + j] = static_cast<int8_t>(i == j ? a : b);
      mat[i * m + m - 1] = 0;
    }
    for (int j = 0; j < m; ++j) mat[(m - 1) * m + j] = 0;
  }
  const int8_t q = static_cast<int8_t>(o.gapOpenPenalty), e = static_cast<int8_t>(o.gapExtendPenalty);
  int w = o.dpBandwidth;
  int32_t mqe = KSW_NEG_INF, mte = KSW_NEG_INF;
  if (qlen <= 0 || tlen <= 0) return std::max(mqe, mte);
  const int qe = q + e;
  const int8_t qe2 = static_cast<int8_t>((q + e) * 2);
  const uint8_t max_sc_b = static_cast<uint8_t>(static_cast<int8_t>(mat[0] + (q + e) * 2));
  const int8_t sc_mch = mat[0], sc_mis = mat[1], sc_N = mat[m * m - 1];
  if (w < 0) w = tlen > qlen ? tlen : qlen;
  const int wl = w, wr = w;
  const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
  int min_sc = mat[1];
  for (int t = 1; t < m * m; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
  if (-min_sc > 2 * (q + e)) return std::max(mqe, mte);

  std::vector<uint8_t> mem(static_cast<size_t>(tlen_ * 6 + qlen_ + 1) * 16, 0);
  uint8_t* u = mem.data();
  uint8_t* v = u + tlen_ * 16;
  uint8_t* x = v + tlen_ * 16;
  uint8_t* y = x + tlen_ * 16;
  uint8_t* s = y + tlen_ * 16;
  uint8_t* sf = s + tlen_ * 16;
  uint8_t* qr = sf + tlen_ * 16;
  std::vector<int32_t> H(static_cast<size_t>(tlen_) * 16, KSW_NEG_INF);
  for (int t = 0; t < qlen; ++t) qr[t] = nt4(qs[qlen - 1 - t]);
  for (int t = 0; t < tlen; ++t) sf[t] = nt4(ts[t]);
  std::vector<uint8_t> xo(static_cast<size_t>(tlen_) * 16 + 16), vo(static_cast<size_t>(tlen_) * 16 + 16);

  int last_st = -1, last_en = -1;
  for (int r = 0; r < qlen + tlen - 1; ++r) {
    int st = 0, en = tlen - 1;
    if (st < r - qlen + 1) st = r - qlen + 1;
    if (en > r) en = r;
    if (st < ((r - wr + 1) >> 1)) st = (r - wr + 1) >> 1;
    if (en > ((r + wl) >> 1)) en = (r + wl) >> 1;
    if (st > en) break; // zdropped
    const int st0 = st, en0 = en;
    ops.kswCells += static_cast<uint64_t>(en0 - st0 + 1);  // cells of the band on this anti-diagonal (before the 16-lane rounding)
    st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
    int8_t x1, v1;
    if (st > 0) {
      if (st - 1 >= last_st && st - 1 <= last_en) { x1 = static_cast<int8_t>(x[st - 1]); v1 = static_cast<int8_t>(v[st - 1]); }
      else x1 = v1 = 0;
    } else { x1 = 0; v1 = r ? q : 0; }
    if (en >= r) { y[r] = 0; u[r] = static_cast<uint8_t>(r ? q : 0); }
    const uint8_t* qrr = qr + (qlen - 1 - r);
    for (int t = st0; t <= en0; t += 16) { // whole 16-lane score blocks (:126-140)
      for (int l = 0; l < 16; ++l) {
        uint8_t sq = sf[t + l], sq2 = qrr[t + l];
        int8_t sc = (sq == sq2) ? sc_mch : sc_mis;
        if (sq == m - 1 || sq2 == m - 1) sc = sc_N;
        s[t + l] = static_cast<uint8_t>(sc);
      }
    }
    // core loop (:147-164): every lane reads previous-row state only
    for (int t = st; t <= en; ++t) { xo[t] = x[t]; vo[t] = v[t]; }
    for (int t = st; t <= en; ++t) {
      int8_t xt1 = (t == st) ? x1 : static_cast<int8_t>(xo[t - 1]);
      int8_t vt1 = (t == st) ? v1 : static_cast<int8_t>(vo[t - 1]);
      int8_t ut = static_cast<int8_t>(u[t]);
      int8_t z = static_cast<int8_t>(static_cast<int8_t>(s[t]) + qe2);
      int8_t a = static_cast<int8_t>(xt1 + vt1);
      int8_t b = static_cast<int8_t>(static_cast<int8_t>(y[t]) + ut);
      z = z > a ? z : a;                                                       // _mm_max_epi8
      uint8_t zu = static_cast<uint8_t>(z), bu = static_cast<uint8_t>(b);
      zu = zu > bu ? zu : bu;                                                  // _mm_max_epu8
      zu = zu < max_sc_b ? zu : max_sc_b;                                      // _mm_min_epu8
      z = static_cast<int8_t>(zu);
      u[t] = static_cast<uint8_t>(static_cast<int8_t>(z - vt1));
      v[t] = static_cast<uint8_t>(static_cast<int8_t>(z - ut));
      z = static_cast<int8_t>(z - q);
      a = static_cast<int8_t>(a - z);
      b = static_cast<int8_t>(b - z);
      x[t] = static_cast<uint8_t>(a > 0 ? a : 0);
      y[t] = static_cast<uint8_t>(b > 0 ? b : 0);
    }
    // exact max (:228-272)
    if (r > 0) {
      H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] - qe : H[en0] + v[en0] - qe;
      for (int t = st0; t < en0; ++t) H[t] += static_cast<int32_t>(v[t]) - qe;
    } else {
      H[0] = v[0] - qe - qe;
    }
    if (en0 == tlen - 1 && H[en0] > mte) mte = H[en0];
    if (r - st0 == qlen - 1 && H[st0] > mqe) mqe = H[st0];
    last_st = st; last_en = en;
  }
  return std::max(mqe, mte);
}

namespace {
struct AlnCache { // tsl::hopscotch_map<uint64_t,int32_t> keyed by MetroHash64 of the window == exact window identity
  std::vector<std::pair<std::string, int32_t>> e;
  bool empty() const { return e.empty(); }
  void clear() { e.clear(); }
  const int32_t* find(const char* p, uint32_t n) const {
    for (auto& kv : e) if (kv.first.size() == n && std::memcmp(kv.first.data(), p, n) == 0) return &kv.second;
    return nullptr;
  }
  void put(const char* p, uint32_t n, int32_t s) {
    for (auto& kv : e) if (kv.first.size() == n && std::memcmp(kv.first.data(), p, n) == 0) { kv.second = s; return; }
    e.emplace_back(std::string(p, n), s);
  }
};

// include/SelectiveAlignmentUtils.hpp:260-373
int32_t getAlnScore(Mapper& M, int32_t pos, const char* rptr, int32_t rlen, const char* tseq, int32_t tlen,
                    int8_t mscore, int8_t mmcost, int32_t maxScore, uint8_t chainStat, bool multiMapping, int ap, uint32_t buf,
                    AlnCache& cache) {
  ++M.ops.alnCalls;
  if (chainStat == PERFECT) return maxScore;
  int32_t s = std::numeric_limits<int32_t>::lowest();
  bool invalidStart = (pos < 0);
  bool invalidEnd = (pos + rlen >= tlen);
  if (invalidStart) { rptr += -pos; rlen += pos; pos = 0; }
  if (invalidStart || invalidEnd) { if (ap == 1 || ap == 2) return s; }
  if (pos < tlen) {
    bool doUngapped = (!invalidStart) && (chainStat == UNGAPPED);
    buf = doUngapped ? 0 : buf;
    uint32_t lnobuf = static_cast<uint32_t>(tlen - pos);
    uint32_t lbuf = static_cast<uint32_t>(rlen + buf);
    bool useBuf = (lbuf < lnobuf);
    uint32_t tlen1 = std::min(lbuf, lnobuf);
    const char* tseq1 = tseq + pos;
    uint32_t keyLen = useBuf ? tlen1 - buf : tlen1;
    if (!cache.empty()) {
      const int32_t* hit = cache.find(tseq1, keyLen);
      if (hit) s = *hit;
    }
    if (s == std::numeric_limits<int32_t>::lowest()) {
      if (doUngapped) {
        int32_t tlen1s = static_cast<int32_t>(tlen1);
        int32_t alnLen = rlen < tlen1s ? rlen : tlen1s;
        int32_t sc = 0;
        for (int32_t i = 0; i < alnLen; ++i) {
          char c1 = tseq1[i], c2 = rptr[i];
          c1 = (c1 == 'N' || c2 == 'N') ? c2 : c1;
          sc += (c1 == c2) ? mscore : mmcost;
        }
        s = sc;
      } else {
        s = M.kswExtzScore(rptr, rlen, tseq1, static_cast<int>(tlen1));
      }
      if (multiMapping) cache.put(tseq1, keyLen, s);
    }
  }
  return s;
}
} // namespace

// edlibAlign(query, target, {k, EDLIB_MODE_HW, EDLIB_TASK_DISTANCE}) as recoverOrphans uses it
// (include/SelectiveAlignmentUtils.hpp:147-148; third-party edlib vendored at src/edlib.cpp:290): the smallest edit
// distance of the WHOLE query against any substring of the target (free end gaps in the target), -1 when it exceeds k,
// and the first (leftmost) 0-based end position in the target that reaches it.  Edlib gets there with Myers'
// bit-vector algorithm and an Ukkonen band; the value it returns is this DP's.
static int semiGlobalDistance(const char* q, int m, const char* t, int n, int k, int& firstEnd) {
  firstEnd = -1;
  if (m <= 0 || n <= 0) return -1;
  std::vector<int> prev(m + 1), cur(m + 1);
  for (int i = 0; i <= m; ++i) prev[i] = i;   // column 0: the query against the empty prefix
  int best = std::numeric_limits<int>::max();
  for (int j = 1; j <= n; ++j) {
    cur[0] = 0;                                // an alignment may start anywhere in the target
    const char tc = t[j - 1];
    for (int i = 1; i <= m; ++i) {
      int d = prev[i - 1] + (q[i - 1] != tc ? 1 : 0);
      if (prev[i] + 1 < d) d = prev[i] + 1;
      if (cur[i - 1] + 1 < d) d = cur[i - 1] + 1;
      cur[i] = d;
    }
    if (cur[m] < best) { best = cur[m]; firstEnd = j - 1; }
    prev.swap(cur);
  }
  if (best > k) { firstEnd = -1; return -1; }
  return best;
}

// selective_alignment::utils::recoverOrphans (include/SelectiveAlignmentUtils.hpp:35-257)
static void recoverOrphans(const Index& idx, const std::string& leftRead, const std::string& rightRead, const std::vector<QuasiAlignment>& leftHits,
                           const std::vector<QuasiAlignment>& rightHits, std::vector<QuasiAlignment>& jointHits) {
  const int32_t l1 = static_cast<int32_t>(leftRead.size()), l2 = static_cast<int32_t>(rightRead.size());
  const int32_t maxDistRight = l2 / 4, maxDistLeft = l1 / 4;
  std::string rc1, rc2;
  bool haveRc1 = false, haveRc2 = false;
  auto recoverSingle = [&](const QuasiAlignment& anchor, bool anchorIsLeft) {
    const uint32_t txp = anchor.tid;
    const char* tseq = idx.seq.data() + idx.txpOffsets[txp];
    const int32_t anchorPos = anchor.allPositions.front();
    const bool anchorFwd = anchor.fwd;
    const int32_t anchorLen = anchorIsLeft ? l1 : l2, otherLen = anchorIsLeft ? l2 : l1;
    const int32_t maxDist = anchorIsLeft ? maxDistRight : maxDistLeft;
    int32_t lpos = anchorIsLeft ? anchorPos : -1, rpos = anchorIsLeft ? -1 : anchorPos;
    const bool lfwd = anchorIsLeft ? anchorFwd : !anchorFwd, rfwd = anchorIsLeft ? !anchorFwd : anchorFwd;
    const std::string& other = anchorIsLeft ? rightRead : leftRead;
    const uint8_t leftChain = anchorIsLeft ? anchor.chainLeft : static_cast<uint8_t>(REGULAR);
    const uint8_t rightChain = anchorIsLeft ? static_cast<uint8_t>(REGULAR) : anchor.chainRight;
    const int32_t refLength = static_cast<int32_t>(idx.txpLens[txp]);
    const char* rptr;
    int32_t startPos, windowLength;
    if (anchorFwd) {  // look downstream for the reverse complement of the other end
      std::string& rc = anchorIsLeft ? rc2 : rc1;
      bool& have = anchorIsLeft ? haveRc2 : haveRc1;
      if (!have) { reverseRead(other, rc); have = true; }
      rptr = rc.data();
      startPos = std::max(0, anchorPos);
      windowLength = std::min(1000, refLength - startPos);
    } else {          // look upstream for the other end as it is
      rptr = other.data();
      const int32_t endPos = std::min(refLength, anchorPos + anchorLen);
      startPos = std::max(0, endPos - 1000);
      windowLength = std::min(1000, endPos);
    }
    int firstEnd;
    const int dist = semiGlobalDistance(rptr, otherLen, tseq + startPos, windowLength, maxDist, firstEnd);
    if (dist > -1) {
      if (anchorIsLeft) rpos = startPos + firstEnd - otherLen; else lpos = startPos + firstEnd - otherLen;
      const int32_t startRead1 = std::max(lpos, 0), startRead2 = std::max(rpos, 0);
      const bool read1First = startRead1 < startRead2;
      const int32_t fragStartPos = read1First ? startRead1 : startRead2;
      const int32_t fragEndPos = read1First ? (startRead2 + l2) : (startRead1 + l1);
      QuasiAlignment qa;
      qa.tid = txp; qa.pos = lpos; qa.fwd = lfwd; qa.readLen = static_cast<uint32_t>(l1);
      qa.fragLen = static_cast<uint32_t>(fragEndPos - fragStartPos); qa.isPaired = true;
      qa.mateLen = static_cast<uint32_t>(otherLen); qa.matePos = rpos; qa.mateIsFwd = rfwd;
      qa.mateStatus = PAIRED_END_PAIRED;
      qa.chainLeft = leftChain; qa.chainRight = rightChain;
      jointHits.push_back(qa);
    }
  };
  size_t li = 0, ri = 0;
  while (li < leftHits.size() && ri < rightHits.size()) {
    const uint32_t lt = leftHits[li].tid, rt = rightHits[ri].tid;
    if (lt < rt) recoverSingle(leftHits[li++], true);
    else if (rt < lt) recoverSingle(rightHits[ri++], false);
    else { std::fprintf(stderr, "recoverOrphans: transcript in common between left and right hits (the reference exits here)\n"); std::exit(1); }
  }
  while (li < leftHits.size()) recoverSingle(leftHits[li++], true);
  while (ri < rightHits.size()) recoverSingle(rightHits[ri++], false);
}

// src/RapMapSAMapper.cpp:461-711
void Mapper::mapPair(const std::string& r1, const std::string& r2, std::vector<QuasiAlignment>& jointHits) {
  jointHits.clear();
  bool tooManyHits = false;
  ++ctr.numReads;
  HitCollectorInfo leftHC, rightHC;
  std::vector<QuasiAlignment> leftHits, rightHits;
  std::string m1 = r1, m2 = r2;
  bool lh = collect(m1, leftHC);
  bool rh = collect(m2, rightHC);
  hitsToMappingsSimple(PAIRED_END_LEFT, leftHC, leftHits);
  hitsToMappingsSimple(PAIRED_END_RIGHT, rightHC, rightHits);
  bool useSmartIntersect = o.fuzzy || o.selAln;
  if (useSmartIntersect) {
    const MergeResult mergeRes = mergeLeftRightHitsFuzzy(lh, rh, leftHits, rightHits, jointHits, o.maxNumHits, tooManyHits, ctr);
    const bool mergeStatusOK = mergeRes == MergeResult::HAD_EMPTY_INTERSECTION || mergeRes == MergeResult::HAD_ONLY_LEFT || mergeRes == MergeResult::HAD_ONLY_RIGHT;
    if (mergeStatusOK && o.recoverOrphans && !tooManyHits && leftHits.size() + rightHits.size() > 0) {  // :498-530
      // the merge "moved" the orphans into jointHits: the reference swaps them back and starts from an empty joint list
      if (mergeRes == MergeResult::HAD_ONLY_LEFT || mergeRes == MergeResult::HAD_ONLY_RIGHT) jointHits.clear();
      recoverOrphans(idx, r1, r2, leftHits, rightHits, jointHits);
    }
  } else mergeLeftRightHits(leftHits, rightHits, jointHits, o.maxNumHits, tooManyHits, ctr);
  if (jointHits.size() > o.maxNumHits) jointHits.clear();
  if (!jointHits.empty() && o.noOrphans) {
    if (jointHits.front().mateStatus != PAIRED_END_PAIRED) jointHits.clear();
  }
  if (o.selAln && !jointHits.empty()) { // :553-683
    AlnCache cacheL, cacheR;
    const int32_t l1 = static_cast<int32_t>(r1.size()), l2 = static_cast<int32_t>(r2.size());
    std::string rc1, rc2;
    bool have1 = false, have2 = false;
    int8_t a = static_cast<int8_t>(o.matchScore), b = static_cast<int8_t>(o.mismatchPenalty);
    int32_t bestScore = std::numeric_limits<int32_t>::lowest();
    std::vector<int32_t> scores(jointHits.size(), bestScore);
    double optFrac = o.minScoreFraction;
    int32_t maxLeftScore = a * l1, maxRightScore = a * l2;
    bool multiMapping = jointHits.size() > 1;
    const uint32_t buf = 20;
    size_t i = 0;
    for (auto& h : jointHits) {
      int32_t score = std::numeric_limits<int32_t>::min();
      const char* tseq = idx.seq.data() + idx.txpOffsets[h.tid];
      const int32_t tlen = static_cast<int32_t>(idx.txpLens[h.tid]);
      if (h.mateStatus == PAIRED_END_PAIRED) {
        if (!h.fwd && !have1) { reverseRead(r1, rc1); have1 = true; }
        if (!h.mateIsFwd && !have2) { reverseRead(r2, rc2); have2 = true; }
        const char* p1 = h.fwd ? r1.data() : rc1.data();
        const char* p2 = h.mateIsFwd ? r2.data() : rc2.data();
        int32_t s1 = getAlnScore(*this, h.pos, p1, l1, tseq, tlen, a, b, maxLeftScore, h.chainLeft, multiMapping, o.alignmentPolicy, buf, cacheL);
        int32_t s2 = getAlnScore(*this, h.matePos, p2, l2, tseq, tlen, a, b, maxRightScore, h.chainRight, multiMapping, o.alignmentPolicy, buf, cacheR);
        if (h.fwd != h.mateIsFwd && o.noDovetail) {
          if (h.fwd && (h.pos > h.matePos)) { s1 = s2 = std::numeric_limits<int32_t>::min(); }
          else if (h.mateIsFwd && (h.matePos > h.pos)) { s1 = s2 = std::numeric_limits<int32_t>::min(); }
        }
        if ((s1 < (optFrac * maxLeftScore)) || (s2 < (optFrac * maxRightScore))) score = std::numeric_limits<int32_t>::min();
        else score = s1 + s2;
      } else if (h.mateStatus == PAIRED_END_LEFT) {
        if (!h.fwd && !have1) { reverseRead(r1, rc1); have1 = true; }
        const char* p = h.fwd ? r1.data() : rc1.data();
        int32_t s = getAlnScore(*this, h.pos, p, l1, tseq, tlen, a, b, maxLeftScore, h.chainLeft, multiMapping, o.alignmentPolicy, buf, cacheL);
        score = (s < (optFrac * maxLeftScore)) ? std::numeric_limits<int32_t>::min() : s;
      } else if (h.mateStatus == PAIRED_END_RIGHT) {
        if (!h.fwd && !have2) { reverseRead(r2, rc2); have2 = true; }
        const char* p = h.fwd ? r2.data() : rc2.data();
        int32_t s = getAlnScore(*this, h.pos, p, l2, tseq, tlen, a, b, maxRightScore, h.chainRight, multiMapping, o.alignmentPolicy, buf, cacheR);
        score = (s < (optFrac * maxRightScore)) ? std::numeric_limits<int32_t>::min() : s;
      }
      bestScore = (score > bestScore) ? score : bestScore;
      scores[i] = score;
      h.score = score;
      ++i;
    }
    if (bestScore > std::numeric_limits<int32_t>::min()) {
      std::vector<QuasiAlignment> kept;
      for (size_t j = 0; j < jointHits.size(); ++j) {
        bool rem = o.hardFilter ? (scores[j] < bestScore) : (scores[j] == std::numeric_limits<int32_t>::min());
        if (!rem) kept.push_back(std::move(jointHits[j]));
      }
      jointHits.swap(kept);
      double bestScoreD = static_cast<double>(bestScore);
      for (auto& qa : jointHits) {
        qa.alnScore = static_cast<int32_t>(qa.score);
        double vv = bestScoreD - qa.score;
        qa.score = o.hardFilter ? -1.0 : std::exp(-vv);
      }
    } else {
      jointHits.clear();
    }
  } else if (o.noDovetail) { // :684-698
    std::vector<QuasiAlignment> kept;
    for (auto& h : jointHits) {
      bool rem = false;
      if (h.fwd != h.mateIsFwd) {
        if (h.fwd && (h.pos > h.matePos)) rem = true;
        else if (h.mateIsFwd && (h.matePos > h.pos)) rem = true;
      }
      if (!rem) kept.push_back(std::move(h));
    }
    jointHits.swap(kept);
  }
  ctr.totHits += jointHits.size();
}

// src/RapMapSAMapper.cpp:156-371 (unmated reads; same operators, MateStatus::SINGLE_END)
void Mapper::mapSingle(const std::string& r, std::vector<QuasiAlignment>& hits) {
  hits.clear();
  ++ctr.numReads;
  HitCollectorInfo hc;
  std::string m = r;
  collect(m, hc);
  hitsToMappingsSimple(SINGLE_END, hc, hits);
  ctr.totHits += hits.size();  // counted before the clear (:241-246)
  if (hits.size() > o.maxNumHits) { hits.clear(); }
  if (o.selAln && !hits.empty()) {
    AlnCache cache;
    const int32_t l1 = static_cast<int32_t>(r.size());
    std::string rc; bool have = false;
    int8_t a = static_cast<int8_t>(o.matchScore), b = static_cast<int8_t>(o.mismatchPenalty);
    int32_t bestScore = std::numeric_limits<int32_t>::lowest();
    std::vector<int32_t> scores(hits.size(), bestScore);
    double optFrac = o.minScoreFraction;
    int32_t maxReadScore = a * l1;
    bool multiMapping = hits.size() > 1;
    size_t i = 0;
    for (auto& h : hits) {
      const char* tseq = idx.seq.data() + idx.txpOffsets[h.tid];
      const int32_t tlen = static_cast<int32_t>(idx.txpLens[h.tid]);
      if (!h.fwd && !have) { reverseRead(r, rc); have = true; }
      const char* p = h.fwd ? r.data() : rc.data();
      int32_t s = getAlnScore(*this, h.pos, p, l1, tseq, tlen, a, b, maxReadScore, h.chainLeft, multiMapping, o.alignmentPolicy, 20, cache);
      int32_t score = (s < (optFrac * maxReadScore)) ? std::numeric_limits<int32_t>::min() : s;
      bestScore = (score > bestScore) ? score : bestScore;
      scores[i++] = score;
      h.score = score;
    }
    if (bestScore > std::numeric_limits<int32_t>::min()) {
      std::vector<QuasiAlignment> kept;
      for (size_t j = 0; j < hits.size(); ++j) {
        bool rem = o.hardFilter ? (scores[j] < bestScore) : (scores[j] == std::numeric_limits<int32_t>::min());
        if (!rem) kept.push_back(std::move(hits[j]));
      }
      hits.swap(kept);
      double bestScoreD = static_cast<double>(bestScore);
      for (auto& qa : hits) { qa.alnScore = static_cast<int32_t>(qa.score); qa.score = o.hardFilter ? -1.0 : std::exp(-(bestScoreD - qa.score)); }
    } else hits.clear();
  }
}

// =================================================================================================
// SAM text — include/RapMapUtils.hpp:95-110,687-810; src/RapMapUtils.cpp:137-196,313-588
// =================================================================================================
namespace {
std::string processReadName(const std::string& name) {
  size_t splitPos = name.find(' ');
  size_t len = name.size();
  if (splitPos < len) len = splitPos; else splitPos = len;
  if (splitPos > 2 && name[splitPos - 2] == '/') len -= 2;
  return name.substr(0, len);
}
void adjustOverhang(int32_t& pos, uint32_t readLen, uint32_t txpLen, std::string& cigar) {
  int32_t sTxpLen = static_cast<int32_t>(txpLen), sReadLen = static_cast<int32_t>(readLen);
  cigar.clear();
  if (pos + sReadLen < 0) { cigar = std::to_string(readLen) + "S"; pos = 0; }
  else if (pos < 0) {
    int32_t matchLen = sReadLen + pos, clipLen = sReadLen - matchLen;
    cigar = std::to_string(clipLen) + "S" + std::to_string(matchLen) + "M";
    pos = 0;
  } else if (pos > sTxpLen) { cigar = std::to_string(readLen) + "S"; }
  else if (pos + sReadLen > sTxpLen) {
    int32_t matchLen = sTxpLen - pos, clipLen = sReadLen - matchLen;
    cigar = std::to_string(matchLen) + "M" + std::to_string(clipLen) + "S";
  } else { cigar = std::to_string(readLen) + "M"; }
}
void getSamFlags(const QuasiAlignment& q, uint16_t& f1, uint16_t& f2) {
  f1 = 0x1; f1 |= q.isPaired ? 0x2 : 0; f2 = f1;
  bool r1Un = q.mateStatus == PAIRED_END_RIGHT, r2Un = q.mateStatus == PAIRED_END_LEFT;
  f1 |= r1Un ? 0x4 : 0; f2 |= r1Un ? 0x8 : 0;
  f2 |= r2Un ? 0x4 : 0; f1 |= r2Un ? 0x8 : 0;
  f1 |= q.fwd ? 0 : 0x10; f1 |= q.mateIsFwd ? 0 : 0x20;
  f2 |= q.mateIsFwd ? 0 : 0x10; f2 |= q.fwd ? 0 : 0x20;
  f1 |= 0x40; f2 |= 0x80;
}
} // namespace

std::string Mapper::samHeader() const {
  std::string h = "@HD\tVN:1.0\tSO:unknown\n";
  for (size_t i = 0; i < idx.txpNames.size(); ++i) h += "@SQ\tSN:" + idx.txpNames[i] + "\tLN:" + std::to_string(idx.txpLens[i]) + "\n";
  h += "@PG\tID:rapmap\tPN:rapmap\tVN:0.6.0\n";
  return h;
}

void Mapper::samPair(const std::string& n1, const std::string& s1, const std::string& n2, const std::string& s2,
                     std::vector<QuasiAlignment>& jointHits, std::string& out) {
  std::string rn = processReadName(n1), mn = processReadName(n2);
  auto T = [](long v) { return std::to_string(v); };
  if (jointHits.empty() || jointHits.size() > o.maxNumHits) { // src/RapMapUtils.cpp:137-196
    out += rn + "\t77\t*\t0\t255\t*\t*\t*\t0\t" + s1 + "\t*\tNH:i:0\tHI:i:0\tAS:i:0\n";
    out += mn + "\t141\t*\t0\t255\t*\t*\t*\t0\t" + s2 + "\t*\tNH:i:0\tHI:i:0\tAS:i:0\n";
    return;
  }
  std::string nh = "NH:i:" + T(static_cast<long>(jointHits.size()));
  std::string rev1, rev2, c1, c2;
  bool haveRev1 = false, haveRev2 = false;
  uint32_t alnCtr = 0;
  size_t i = 0;
  for (auto& qa : jointHits) {
    ++i;
    const std::string& tn = idx.txpNames[qa.tid];
    uint32_t txpLen = static_cast<uint32_t>(idx.txpLens[qa.tid]);
    uint16_t f1, f2;
    getSamFlags(qa, f1, f2);
    if (alnCtr != 0) { f1 |= 0x100; f2 |= 0x100; }
    std::string tail = "\t*\t" + nh + "\tHI:i:" + T(static_cast<long>(i)) + "\tAS:i:" + T(qa.alnScore) + "\n";
    if (qa.isPaired) {
      adjustOverhang(qa.pos, qa.readLen, txpLen, c1);
      adjustOverhang(qa.matePos, qa.mateLen, txpLen, c2);
      const std::string* q1 = &s1;
      if (!qa.fwd) { if (!haveRev1) { reverseRead(s1, rev1); haveRev1 = true; } q1 = &rev1; }
      const std::string* q2 = &s2;
      if (!qa.mateIsFwd) { if (!haveRev2) { reverseRead(s2, rev2); haveRev2 = true; } q2 = &rev2; }
      int32_t p1 = qa.pos, p2 = qa.matePos;
      bool read1First = p1 < p2;
      int32_t minPos = read1First ? p1 : p2;
      if ((minPos + static_cast<int32_t>(qa.fragLen)) > static_cast<int32_t>(txpLen)) qa.fragLen = txpLen - minPos;
      int32_t fragLen = static_cast<int32_t>(qa.fragLen);
      out += rn + "\t" + T(f1) + "\t" + tn + "\t" + T(qa.pos + 1) + "\t1\t" + c1 + "\t=\t" + T(qa.matePos + 1) + "\t" +
             T(read1First ? fragLen : -fragLen) + "\t" + *q1 + tail;
      out += mn + "\t" + T(f2) + "\t" + tn + "\t" + T(qa.matePos + 1) + "\t1\t" + c2 + "\t=\t" + T(qa.pos + 1) + "\t" +
             T(read1First ? -fragLen : fragLen) + "\t" + *q2 + tail;
    } else {
      bool left = qa.mateStatus == PAIRED_END_LEFT;
      const std::string& an = left ? rn : mn;
      const std::string& un = left ? mn : rn;
      const std::string* rs = left ? &s1 : &s2;
      const std::string& us = left ? s2 : s1;
      uint32_t fl = left ? f1 : f2, ufl = left ? f2 : f1;
      std::string& cg = left ? c1 : c2;
      if (!qa.fwd) {
        bool& have = left ? haveRev1 : haveRev2;
        std::string& tmp = left ? rev1 : rev2;
        if (!have) { reverseRead(*rs, tmp); have = true; }
        rs = &tmp;
      }
      adjustOverhang(qa.pos, qa.readLen, txpLen, cg);
      out += an + "\t" + T(fl) + "\t" + tn + "\t" + T(qa.pos + 1) + "\t1\t" + cg + "\t=\t" + T(qa.pos + 1) + "\t0\t" + *rs + tail;
      out += un + "\t" + T(ufl) + "\t" + tn + "\t" + T(qa.pos + 1) + "\t0\t*\t=\t" + T(qa.pos + 1) + "\t0\t" + us + tail;
    }
    ++alnCtr;
  }
}

// src/RapMapUtils.cpp:198-311 (unmated reads): writeUnalignedSingleToStream / the single-read writeAlignmentsToStream with
// getSamFlags(qa, flags) of include/RapMapUtils.hpp:737-768.  The name is cut at the first space only.
void Mapper::samSingle(const std::string& name, const std::string& seq, std::vector<QuasiAlignment>& hits, std::string& out) {
  std::string rn = name.substr(0, std::min(name.find(' '), name.size()));
  auto T = [](long v) { return std::to_string(v); };
  if (hits.empty()) {
    out += rn + "\t4\t*\t0\t255\t*\t*\t0\t0\t" + seq + "\t*\tNH:i:0\tHI:i:0\tAS:i:0\n";
    return;
  }
  std::string nh = "NH:i:" + T(static_cast<long>(hits.size()));
  std::string rev, cigar;
  bool haveRev = false;
  uint32_t alnCtr = 0;
  size_t i = 0;
  for (auto& qa : hits) {
    ++i;
    uint16_t flags = qa.fwd ? 0 : 0x10;
    if (alnCtr != 0) flags |= 0x900;
    const std::string* rs = &seq;
    if (!qa.fwd) { if (!haveRev) { reverseRead(seq, rev); haveRev = true; } rs = &rev; }
    adjustOverhang(qa.pos, qa.readLen, static_cast<uint32_t>(idx.txpLens[qa.tid]), cigar);
    out += rn + "\t" + T(flags) + "\t" + idx.txpNames[qa.tid] + "\t" + T(qa.pos + 1) + "\t255\t" + cigar + "\t*\t0\t" + T(qa.fragLen) + "\t" + *rs +
           "\t*\t" + nh + "\tHI:i:" + T(static_cast<long>(i)) + "\tAS:i:" + T(qa.alnScore) + "\n";
    ++alnCtr;
  }
}

} // namespace oracle
