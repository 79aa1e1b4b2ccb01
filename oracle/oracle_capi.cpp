// CPU ORACLE — TEST INFRASTRUCTURE ONLY (see quasimap_oracle.hpp).
// C entry points (ctypes) and a `quasimap`-like CLI around the restatement, so the tests can compare
// (a) oracle SAM with the compiled reference's SAM byte for byte and (b) the CUDA path's
// rapmap_hit_t records with the oracle's on the same inputs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../include/rapmap_cuda.h"
#include "quasimap_oracle.hpp"

using namespace oracle;

static Opts fromC(const rapmap_cuda_opts_t* c) {
  Opts o;
  o.maxNumHits = c->max_num_hits; o.quasiCov = c->quasi_coverage; o.sensitive = c->sensitive; o.strictCheck = c->strict_check;
  o.fuzzy = c->fuzzy; o.selAln = c->sel_aln; o.consensusSlack = c->consensus_slack; o.minScoreFraction = c->min_score_fraction;
  o.matchScore = c->match_score; o.mismatchPenalty = c->mismatch_penalty; o.gapOpenPenalty = c->gap_open_penalty;
  o.gapExtendPenalty = c->gap_extend_penalty; o.dpBandwidth = c->dp_bandwidth; o.hardFilter = c->hard_filter;
  o.alignmentPolicy = c->alignment_policy; o.noOrphans = c->no_orphans; o.noDovetail = c->no_dovetail;
  o.maxMMPExtension = c->max_mmp_extension; o.recoverOrphans = c->recover_orphans;
  return o;
}

static void toHit(const QuasiAlignment& q, rapmap_hit_t& h) {
  std::memset(&h, 0, sizeof(h));
  h.tid = q.tid; h.pos = q.pos; h.read_len = static_cast<uint16_t>(q.readLen); h.fwd = q.fwd; h.mate_fwd = q.mateIsFwd;
  h.mate_status = q.mateStatus; h.aln_score = q.alnScore; h.chain_status = static_cast<uint8_t>(q.chainLeft | (q.chainRight << 4));
  if (q.mateStatus == PAIRED_END_PAIRED) { h.mate_pos = q.matePos; h.mate_len = static_cast<uint16_t>(q.mateLen); h.frag_len = q.fragLen; }
}

extern "C" {

void* oracle_index_load(const char* dir, char* err, int errlen) {
  auto* idx = new Index();
  std::string e;
  if (!idx->load(dir, &e)) {
    if (err && errlen > 0) { std::strncpy(err, e.c_str(), static_cast<size_t>(errlen) - 1); err[errlen - 1] = 0; }
    delete idx;
    return nullptr;
  }
  return idx;
}
void oracle_index_free(void* p) { delete static_cast<Index*>(p); }
uint64_t oracle_index_num_kmers(void* p) { return static_cast<Index*>(p)->khash.size(); }
uint64_t oracle_index_text_len(void* p) { return static_cast<Index*>(p)->seq.size(); }

void* oracle_mapper_new(void* idx, const rapmap_cuda_opts_t* o) { return new Mapper(*static_cast<Index*>(idx), fromC(o)); }
void oracle_mapper_free(void* p) { delete static_cast<Mapper*>(p); }

static std::string getRead(const uint8_t* seq, const uint64_t* off, uint32_t fixedLen, uint64_t i) {
  if (off) return std::string(reinterpret_cast<const char*>(seq) + off[i], off[i + 1] - off[i]);
  return std::string(reinterpret_cast<const char*>(seq) + i * fixedLen, fixedLen);
}

// Same contract as rapmap_cuda_map_batch (host buffers only). Returns 0, or 5 if capacity is too small.
int oracle_map_batch(void* mp, const rapmap_read_batch_t* reads, rapmap_hit_batch_t* out) {
  Mapper& M = *static_cast<Mapper*>(mp);
  M.ctr = Counters();
  std::vector<QuasiAlignment> joint;
  uint64_t nh = 0;
  bool overflow = false;
  for (uint64_t i = 0; i < reads->n; ++i) {
    out->pair_offsets[i] = nh;
    std::string r1 = getRead(reads->seq1, reads->off1, reads->fixed_len, i);
    if (reads->seq2) {
      std::string r2 = getRead(reads->seq2, reads->off2, reads->fixed_len, i);
      M.mapPair(r1, r2, joint);
    } else {
      M.mapSingle(r1, joint);
    }
    for (auto& q : joint) {
      if (nh < out->hits_capacity) toHit(q, out->hits[nh]); else overflow = true;
      ++nh;
    }
  }
  out->pair_offsets[reads->n] = nh;
  out->num_hits = nh;
  out->counters[0] = M.ctr.numReads; out->counters[1] = M.ctr.peHits; out->counters[2] = M.ctr.seHits;
  out->counters[3] = M.ctr.totHits; out->counters[4] = M.ctr.tooManyHits;
  return overflow ? 5 : 0;
}

// Stage tap: SAIntervalHit lists of one read (SACollector::operator()).
int oracle_collect(void* mp, const uint8_t* seq, uint32_t len, rapmap_sa_interval_t* out, uint32_t cap, uint32_t* nFwd, uint32_t* nRc,
                   uint8_t* found) {
  Mapper& M = *static_cast<Mapper*>(mp);
  std::string r(reinterpret_cast<const char*>(seq), len);
  HitCollectorInfo hc;
  *found = M.collect(r, hc);
  *nFwd = static_cast<uint32_t>(hc.fwdSAInts.size());
  *nRc = static_cast<uint32_t>(hc.rcSAInts.size());
  uint32_t w = 0;
  for (auto* v : {&hc.fwdSAInts, &hc.rcSAInts})
    for (auto& h : *v) {
      if (w < cap) { out[w].begin = h.begin; out[w].end = h.end; out[w].len = h.len; out[w].query_pos = h.queryPos; out[w].query_rc = h.queryRC; }
      ++w;
    }
  return w <= cap ? 0 : 5;
}

// Operation counters accumulated since the mapper was created: hashFind, saProbes,
// textCmp, rankCalls, intervals, kswCalls, alnCalls, kswCells (8 values).
void oracle_op_counts(void* mp, uint64_t* out7) {
  Mapper& M = *static_cast<Mapper*>(mp);
  out7[0] = M.ops.hashFind; out7[1] = M.ops.saProbes; out7[2] = M.ops.textCmp; out7[3] = M.ops.rankCalls;
  out7[4] = M.ops.intervals; out7[5] = M.ops.kswCalls; out7[6] = M.ops.alnCalls; out7[7] = M.ops.kswCells;
}

// Known-answer hook for the DP alone.
int32_t oracle_ksw_extz_score(void* mp, const char* q, int qlen, const char* t, int tlen) {
  return static_cast<Mapper*>(mp)->kswExtzScore(q, qlen, t, tlen);
}

} // extern "C"

#ifdef ORACLE_MAIN
// quasimap_oracle -i <index> -1 r1.fastq -2 r2.fastq [-r reads.fastq] [-s] [-o out.sam] [flags of src/RapMapSAMapper.cpp:992-1023]
static bool nextFastq(std::ifstream& f, std::string& name, std::string& seq) {
  std::string l, plus, qual;
  if (!std::getline(f, l)) return false;
  if (l.empty()) return false;
  name = l.substr(1);
  // kseq: name = up to first whitespace (include/kseq.h); the SAM writer splits on ' ' again
  size_t ws = name.find_first_of(" \t");
  if (ws != std::string::npos) name.resize(ws);
  if (!std::getline(f, seq)) return false;
  if (l[0] == '@') { std::getline(f, plus); std::getline(f, qual); }
  return true;
}

int main(int argc, char** argv) {
  rapmap_cuda_opts_t c;
  std::memset(&c, 0, sizeof(c));
  c.max_num_hits = 200; c.sensitive = 1; c.strict_check = 1; c.consensus_slack = 0.2f; c.min_score_fraction = 0.65;
  c.match_score = 2; c.mismatch_penalty = -4; c.gap_open_penalty = 4; c.gap_extend_penalty = 2; c.dp_bandwidth = 15; c.max_mmp_extension = 7;
  std::string index, r1, r2, ru, outname;
  bool mimicBT2 = false, mimicStrict = false, printOps = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
    if (a == "-i" || a == "--index") index = val();
    else if (a == "-1" || a == "--leftMates") r1 = val();
    else if (a == "-2" || a == "--rightMates") r2 = val();
    else if (a == "-r" || a == "--unmatedReads") ru = val();
    else if (a == "-o" || a == "--output") outname = val();
    else if (a == "-t" || a == "--numThreads") val();
    else if (a == "-m" || a == "--maxNumHits") c.max_num_hits = static_cast<uint32_t>(std::stoul(val()));
    else if (a == "-z" || a == "--quasiCoverage") c.quasi_coverage = std::stod(val());
    else if (a == "--noSensitive") c.sensitive = 0;
    else if (a == "--noStrictCheck") c.strict_check = 0;
    else if (a == "-f" || a == "--fuzzyIntersection") c.fuzzy = 1;
    else if (a == "-c" || a == "--chaining") {}
    else if (a == "-s" || a == "--selAln") c.sel_aln = 1;
    else if (a == "--go") c.gap_open_penalty = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--ge") c.gap_extend_penalty = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--mm") c.mismatch_penalty = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--ma") c.match_score = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--dpBandwidth") c.dp_bandwidth = std::stoi(val());
    else if (a == "--minScoreFrac") c.min_score_fraction = std::stod(val());
    else if (a == "--consensusSlack") c.consensus_slack = std::stof(val());
    else if (a == "--hardFilter") c.hard_filter = 1;
    else if (a == "--noOrphans") c.no_orphans = 1;
    else if (a == "--noDovetail") c.no_dovetail = 1;
    else if (a == "--mimicBT2") mimicBT2 = true;
    else if (a == "--mimicStrictBT2") mimicStrict = true;
    else if (a == "--maxMMPExtension") c.max_mmp_extension = std::stoi(val());
    else if (a == "--recoverOrphans") c.recover_orphans = 1;
    else if (a == "--ops") printOps = true;
    else if (a == "-n" || a == "--noOutput" || a == "-q" || a == "--quiet") {}
    else { std::fprintf(stderr, "unknown flag %s\n", a.c_str()); return 1; }
  }
  // src/RapMapSAMapper.cpp:1119-1174
  if ((mimicBT2 || mimicStrict) && !c.sel_aln) c.sel_aln = 1;
  if (c.sel_aln) {
    if (mimicBT2) { c.alignment_policy = 1; c.no_orphans = 1; c.no_dovetail = 1; c.consensus_slack = 0.35; c.max_num_hits = 1000; }
    if (mimicStrict) {
      c.alignment_policy = 2; c.no_orphans = 1; c.no_dovetail = 1; c.consensus_slack = 0.35; c.max_num_hits = 1000;
      c.min_score_fraction = 0.8; c.match_score = 1; c.mismatch_penalty = 0; c.gap_open_penalty = 25; c.gap_extend_penalty = 25;
    }
  }
  if (c.quasi_coverage > 0 && !c.sensitive) c.sensitive = 1;
  Index idx;
  std::string err;
  if (!idx.load(index, &err)) { std::fprintf(stderr, "index load failed: %s\n", err.c_str()); return 1; }
  Mapper M(idx, fromC(&c));
  FILE* out = outname.empty() ? stdout : std::fopen(outname.c_str(), "w");
  std::string hdr = M.samHeader();
  std::fwrite(hdr.data(), 1, hdr.size(), out);
  std::vector<QuasiAlignment> joint;
  std::string sam;
  if (!r1.empty()) {
    std::ifstream f1(r1), f2(r2);
    std::string n1, s1, n2, s2;
    while (nextFastq(f1, n1, s1) && nextFastq(f2, n2, s2)) {
      M.mapPair(s1, s2, joint);
      sam.clear();
      M.samPair(n1, s1, n2, s2, joint, sam);
      std::fwrite(sam.data(), 1, sam.size(), out);
    }
  } else if (!ru.empty()) {
    std::ifstream f(ru);
    std::string n, sq;
    while (nextFastq(f, n, sq)) {
      M.mapSingle(sq, joint);
      sam.clear();
      M.samSingle(n, sq, joint, sam);
      std::fwrite(sam.data(), 1, sam.size(), out);
    }
  } else {
    std::fprintf(stderr, "no reads given (-1/-2 or -r)\n");
    return 1;
  }
  if (out != stdout) std::fclose(out);
  std::fprintf(stderr, "oracle: reads %llu  hits/read %.5f\n", static_cast<unsigned long long>(M.ctr.numReads),
               M.ctr.totHits / static_cast<float>(M.ctr.numReads));
  if (printOps) {
    double n = static_cast<double>(M.ctr.numReads);
    std::fprintf(stderr, "ops/pair: hashFind %.2f saProbes %.2f textCmp %.2f rank %.2f intervals %.2f ksw %.3f aln %.3f\n",
                 M.ops.hashFind / n, M.ops.saProbes / n, M.ops.textCmp / n, M.ops.rankCalls / n, M.ops.intervals / n,
                 M.ops.kswCalls / n, M.ops.alnCalls / n);
  }
  return 0;
}
#endif
