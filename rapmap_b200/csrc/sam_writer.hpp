// Host-side SAM text for a chunk of mapped reads.
// Follows the reference's writers field by field so `quasimap` output stays byte-identical:
//   writeAlignmentsToStream (paired)  src/RapMapUtils.cpp:313-588
//   writeUnalignedPairToStream        src/RapMapUtils.cpp:137-196
//   adjustOverhang / getSamFlags      include/RapMapUtils.hpp:687-810
//   writeSAMHeader                    include/RapMapUtils.hpp:95-110 (VN = rapmap::version 0.6.0)
// SAM formatting is row f1 of SURVEY.md §8 ("next"): it stays on the host, fed by rapmap_hit_t records.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rapmap_cuda.h"

namespace rapmap_b200 {

inline void samReverseRead(const char* s, size_t n, std::string& out) {  // rapmap::utils::reverseRead
  out.resize(n);
  for (size_t i = 0; i < n; ++i) {
    char r;
    switch (s[n - 1 - i]) {
      case 'A': case 'a': r = 'T'; break;
      case 'C': case 'c': r = 'G'; break;
      case 'G': case 'g': r = 'C'; break;
      case 'T': case 't': case 'U': case 'u': r = 'A'; break;
      default: r = 'N';
    }
    out[i] = r;
  }
}

inline void samReadName(const char* name, std::string& out) {  // processReadName lambda
  size_t len = std::char_traits<char>::length(name);
  size_t split = len;
  for (size_t i = 0; i < len; ++i)
    if (name[i] == ' ') { split = i; break; }
  size_t keep = split;
  if (split > 2 && name[split - 2] == '/') keep -= 2;
  out.assign(name, keep);
}

inline void samOverhang(int32_t& pos, uint32_t readLen, uint32_t txpLen, std::string& cigar) {
  const int32_t sT = static_cast<int32_t>(txpLen), sR = static_cast<int32_t>(readLen);
  cigar.clear();
  if (pos + sR < 0) {
    cigar = std::to_string(readLen) + "S";
    pos = 0;
  } else if (pos < 0) {
    int32_t matchLen = sR + pos, clipLen = sR - matchLen;
    cigar = std::to_string(clipLen) + "S" + std::to_string(matchLen) + "M";
    pos = 0;
  } else if (pos > sT) {
    cigar = std::to_string(readLen) + "S";
  } else if (pos + sR > sT) {
    int32_t matchLen = sT - pos, clipLen = sR - matchLen;
    cigar = std::to_string(matchLen) + "M" + std::to_string(clipLen) + "S";
  } else {
    cigar = std::to_string(readLen) + "M";
  }
}

inline void samAppendInt(std::string& o, long v) { o += std::to_string(v); }

inline std::string samHeader(const std::vector<std::string>& names, const std::vector<int32_t>& lens) {
  std::string h = "@HD\tVN:1.0\tSO:unknown\n";
  for (size_t i = 0; i < names.size(); ++i) {
    h += "@SQ\tSN:";
    h += names[i];
    h += "\tLN:";
    samAppendInt(h, lens[i]);
    h += '\n';
  }
  h += "@PG\tID:rapmap\tPN:rapmap\tVN:0.6.0\n";
  return h;
}

// One read pair.  `hits` is modified the way the reference modifies jointHits while printing (clamped
// positions, fragLen clipped to the transcript end).
inline void samPair(const std::vector<std::string>& names, const std::vector<int32_t>& lens, uint32_t maxNumHits, const char* name1,
                    const char* s1, size_t l1, const char* name2, const char* s2, size_t l2, rapmap_hit_t* hits, size_t nh, std::string& out) {
  std::string rn, mn;
  samReadName(name1, rn);
  samReadName(name2, mn);
  if (nh == 0 || nh > maxNumHits) {
    out += rn; out += "\t77\t*\t0\t255\t*\t*\t*\t0\t"; out.append(s1, l1); out += "\t*\tNH:i:0\tHI:i:0\tAS:i:0\n";
    out += mn; out += "\t141\t*\t0\t255\t*\t*\t*\t0\t"; out.append(s2, l2); out += "\t*\tNH:i:0\tHI:i:0\tAS:i:0\n";
    return;
  }
  std::string nhFlag = "NH:i:" + std::to_string(nh);
  std::string rev1, rev2, c1, c2, tail;
  bool haveRev1 = false, haveRev2 = false;
  for (size_t i = 0; i < nh; ++i) {
    rapmap_hit_t& qa = hits[i];
    const std::string& tn = names[qa.tid];
    const uint32_t txpLen = static_cast<uint32_t>(lens[qa.tid]);
    const bool isPaired = qa.mate_status == 3;
    uint16_t f1 = 0x1 | (isPaired ? 0x2 : 0), f2 = f1;
    const bool r1Un = qa.mate_status == 2, r2Un = qa.mate_status == 1;
    f1 |= r1Un ? 0x4 : 0; f2 |= r1Un ? 0x8 : 0;
    f2 |= r2Un ? 0x4 : 0; f1 |= r2Un ? 0x8 : 0;
    f1 |= qa.fwd ? 0 : 0x10; f1 |= qa.mate_fwd ? 0 : 0x20;
    f2 |= qa.mate_fwd ? 0 : 0x10; f2 |= qa.fwd ? 0 : 0x20;
    f1 |= 0x40; f2 |= 0x80;
    if (i != 0) { f1 |= 0x100; f2 |= 0x100; }
    tail = "\t*\t" + nhFlag + "\tHI:i:" + std::to_string(i + 1) + "\tAS:i:" + std::to_string(qa.aln_score) + "\n";
    if (isPaired) {
      samOverhang(qa.pos, qa.read_len, txpLen, c1);
      samOverhang(qa.mate_pos, qa.mate_len, txpLen, c2);
      const char* q1 = s1; size_t q1l = l1;
      if (!qa.fwd) { if (!haveRev1) { samReverseRead(s1, l1, rev1); haveRev1 = true; } q1 = rev1.data(); q1l = rev1.size(); }
      const char* q2 = s2; size_t q2l = l2;
      if (!qa.mate_fwd) { if (!haveRev2) { samReverseRead(s2, l2, rev2); haveRev2 = true; } q2 = rev2.data(); q2l = rev2.size(); }
      const bool read1First = qa.pos < qa.mate_pos;
      const int32_t minPos = read1First ? qa.pos : qa.mate_pos;
      if ((minPos + static_cast<int32_t>(qa.frag_len)) > static_cast<int32_t>(txpLen)) qa.frag_len = txpLen - minPos;
      const int32_t fragLen = static_cast<int32_t>(qa.frag_len);
      out += rn; out += '\t'; samAppendInt(out, f1); out += '\t'; out += tn; out += '\t'; samAppendInt(out, qa.pos + 1); out += "\t1\t"; out += c1;
      out += "\t=\t"; samAppendInt(out, qa.mate_pos + 1); out += '\t'; samAppendInt(out, read1First ? fragLen : -fragLen); out += '\t';
      out.append(q1, q1l); out += tail;
      out += mn; out += '\t'; samAppendInt(out, f2); out += '\t'; out += tn; out += '\t'; samAppendInt(out, qa.mate_pos + 1); out += "\t1\t"; out += c2;
      out += "\t=\t"; samAppendInt(out, qa.pos + 1); out += '\t'; samAppendInt(out, read1First ? -fragLen : fragLen); out += '\t';
      out.append(q2, q2l); out += tail;
    } else {
      const bool left = qa.mate_status == 1;
      const std::string& an = left ? rn : mn;
      const std::string& un = left ? mn : rn;
      const char* rs = left ? s1 : s2; size_t rl = left ? l1 : l2;
      const char* us = left ? s2 : s1; size_t ul = left ? l2 : l1;
      const uint16_t fl = left ? f1 : f2, ufl = left ? f2 : f1;
      std::string& cg = left ? c1 : c2;
      if (!qa.fwd) {
        bool& have = left ? haveRev1 : haveRev2;
        std::string& tmp = left ? rev1 : rev2;
        if (!have) { samReverseRead(rs, rl, tmp); have = true; }
        rs = tmp.data(); rl = tmp.size();
      }
      samOverhang(qa.pos, qa.read_len, txpLen, cg);
      out += an; out += '\t'; samAppendInt(out, fl); out += '\t'; out += tn; out += '\t'; samAppendInt(out, qa.pos + 1); out += "\t1\t"; out += cg;
      out += "\t=\t"; samAppendInt(out, qa.pos + 1); out += "\t0\t"; out.append(rs, rl); out += tail;
      out += un; out += '\t'; samAppendInt(out, ufl); out += '\t'; out += tn; out += '\t'; samAppendInt(out, qa.pos + 1); out += "\t0\t*\t=\t";
      samAppendInt(out, qa.pos + 1); out += "\t0\t"; out.append(us, ul); out += tail;
    }
  }
}

// One unmated read (processReadsSingleSA, src/RapMapSAMapper.cpp:156-371): writeAlignmentsToStream for single reads
// (src/RapMapUtils.cpp:230-311; flags of getSamFlags(qa, flags), include/RapMapUtils.hpp:737-768: only 0x10, and 0x900
// on every record after the first) or writeUnalignedSingleToStream (src/RapMapUtils.cpp:198-227).  The read name is cut
// at the first space only (no "/1" stripping here).
inline void samSingle(const std::vector<std::string>& names, const std::vector<int32_t>& lens, const char* name, const char* s, size_t l,
                      rapmap_hit_t* hits, size_t nh, std::string& out) {
  size_t nameLen = std::char_traits<char>::length(name);
  for (size_t i = 0; i < nameLen; ++i)
    if (name[i] == ' ') { nameLen = i; break; }
  if (nh == 0) {
    out.append(name, nameLen); out += "\t4\t*\t0\t255\t*\t*\t0\t0\t"; out.append(s, l); out += "\t*\tNH:i:0\tHI:i:0\tAS:i:0\n";
    return;
  }
  const std::string nhFlag = "NH:i:" + std::to_string(nh);
  std::string rev, cigar;
  bool haveRev = false;
  for (size_t i = 0; i < nh; ++i) {
    rapmap_hit_t& qa = hits[i];
    uint16_t flags = qa.fwd ? 0 : 0x10;
    if (i != 0) flags |= 0x900;
    const char* q = s; size_t ql = l;
    if (!qa.fwd) { if (!haveRev) { samReverseRead(s, l, rev); haveRev = true; } q = rev.data(); ql = rev.size(); }
    samOverhang(qa.pos, qa.read_len, static_cast<uint32_t>(lens[qa.tid]), cigar);
    out.append(name, nameLen); out += '\t'; samAppendInt(out, flags); out += '\t'; out += names[qa.tid]; out += '\t'; samAppendInt(out, qa.pos + 1);
    out += "\t255\t"; out += cigar; out += "\t*\t0\t"; samAppendInt(out, qa.frag_len); out += '\t'; out.append(q, ql);
    out += "\t*\t"; out += nhFlag; out += "\tHI:i:"; samAppendInt(out, static_cast<long>(i + 1)); out += "\tAS:i:"; samAppendInt(out, qa.aln_score); out += '\n';
  }
}

} // namespace rapmap_b200
