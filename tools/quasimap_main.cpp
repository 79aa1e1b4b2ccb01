// `rapmap_b200 quasimap` — host front end keeping the reference CLI (flag table: reference
// src/RapMapSAMapper.cpp:992-1023) on top of the C-ABI.  What the reference does per 10,000-read chunk inside
// processReadsPairSA (src/RapMapSAMapper.cpp:461-746) happens here per batch: FASTQ parsing on the host
// (fastx_parser equivalent, include/FastxParser.hpp), one rapmap_cuda_map_batch call, SAM text from the
// returned QuasiAlignment records (rapmap_cuda_format_sam).  Output is byte-identical to `rapmap quasimap -t 1`.
#include <zlib.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/rapmap_cuda.h"

namespace {

struct Fastx {
  gzFile f{nullptr};
  std::vector<char> buf;
  size_t pos{0}, len{0};
  bool open(const std::string& p) {
    f = gzopen(p.c_str(), "rb");
    buf.resize(1 << 20);
    return f != nullptr;
  }
  bool getline(std::string& out) {
    out.clear();
    while (true) {
      if (pos == len) {
        int n = gzread(f, buf.data(), static_cast<unsigned>(buf.size()));
        if (n <= 0) return !out.empty();
        len = static_cast<size_t>(n);
        pos = 0;
      }
      char* b = buf.data() + pos;
      char* e = static_cast<char*>(memchr(b, '\n', len - pos));
      if (e) {
        out.append(b, e - b);
        pos += static_cast<size_t>(e - b) + 1;
        if (!out.empty() && out.back() == '\r') out.pop_back();
        return true;
      }
      out.append(b, len - pos);
      pos = len;
    }
  }
  // kseq semantics: name = header up to the first whitespace
  bool next(std::string& name, std::string& seq) {
    std::string l, tmp;
    do {
      if (!getline(l)) return false;
    } while (l.empty());
    bool fq = l[0] == '@';
    name = l.substr(1);
    size_t ws = name.find_first_of(" \t");
    if (ws != std::string::npos) name.resize(ws);
    if (!getline(seq)) return false;
    if (fq) { getline(tmp); getline(tmp); }
    return true;
  }
  ~Fastx() { if (f) gzclose(f); }
};

void die(const char* what) {
  std::fprintf(stderr, "[rapmap_b200] %s: %s\n", what, rapmap_cuda_last_error());
  std::exit(1);
}

void usage() {
  std::fprintf(stderr,
               "rapmap_b200 quasimap -i <index> -1 <mates1> -2 <mates2> [-o out.sam] [-s] [-m maxNumHits] [-z cov] [-f] [-n]\n"
               "        [--noOrphans] [--noDovetail] [--hardFilter] [--go N --ge N --mm N --ma N] [--dpBandwidth N] [--minScoreFrac F]\n"
               "        [--consensusSlack F] [--maxMMPExtension N] [--mimicBT2 | --mimicStrictBT2] [--device D] [--batch PAIRS]\n");
}

} // namespace

int main(int argc, char** argv) {
  if (argc < 2 || std::strcmp(argv[1], "quasimap") != 0) {
    usage();
    return 1;
  }
  rapmap_cuda_opts_t o;
  rapmap_cuda_opts_default(&o);
  std::string index, r1, r2, outname;
  bool noOutput = false, bt2 = false, strictBt2 = false, quiet = false;
  int device = 0;
  uint64_t batch = 1 << 17;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> std::string { if (i + 1 >= argc) { usage(); std::exit(1); } return argv[++i]; };
    if (a == "-i" || a == "--index") index = val();
    else if (a == "-1" || a == "--leftMates") r1 = val();
    else if (a == "-2" || a == "--rightMates") r2 = val();
    else if (a == "-o" || a == "--output") outname = val();
    else if (a == "-t" || a == "--numThreads") val();  // one GPU stream replaces the worker threads
    else if (a == "-m" || a == "--maxNumHits") o.max_num_hits = static_cast<uint32_t>(std::stoul(val()));
    else if (a == "-z" || a == "--quasiCoverage") o.quasi_coverage = std::stod(val());
    else if (a == "-n" || a == "--noOutput") noOutput = true;
    else if (a == "-q" || a == "--quiet") quiet = true;
    else if (a == "--noSensitive") o.sensitive = 0;
    else if (a == "--noStrictCheck") o.strict_check = 0;
    else if (a == "-f" || a == "--fuzzyIntersection") o.fuzzy = 1;
    else if (a == "-c" || a == "--chaining") {}
    else if (a == "-u" || a == "--writeUnmapped") {}
    else if (a == "-s" || a == "--selAln") o.sel_aln = 1;
    else if (a == "--recoverOrphans") o.recover_orphans = 1;
    else if (a == "--noDovetail") o.no_dovetail = 1;
    else if (a == "--noOrphans") o.no_orphans = 1;
    else if (a == "--hardFilter") o.hard_filter = 1;
    else if (a == "--go") o.gap_open_penalty = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--ge") o.gap_extend_penalty = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--mm") o.mismatch_penalty = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--ma") o.match_score = static_cast<int16_t>(std::stoi(val()));
    else if (a == "--dpBandwidth") o.dp_bandwidth = std::stoi(val());
    else if (a == "--minScoreFrac") o.min_score_fraction = std::stod(val());
    else if (a == "--consensusSlack") o.consensus_slack = std::stof(val());
    else if (a == "--maxMMPExtension") o.max_mmp_extension = std::stoi(val());
    else if (a == "--mimicBT2") bt2 = true;
    else if (a == "--mimicStrictBT2") strictBt2 = true;
    else if (a == "--device") device = std::stoi(val());
    else if (a == "--batch") batch = std::stoull(val());
    else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); usage(); return 1; }
  }
  // option derivation of rapMapSAMap (src/RapMapSAMapper.cpp:1119-1174)
  if (bt2 && strictBt2) { std::fprintf(stderr, "Cannot set --mimicBT2 and --mimicStrictBT2 simultaneously.\n"); return 1; }
  if (bt2 || strictBt2) {
    o.sel_aln = 1;
    o.alignment_policy = bt2 ? 1 : 2;
    o.no_orphans = 1; o.no_dovetail = 1; o.consensus_slack = 0.35; o.max_num_hits = 1000;
    if (strictBt2) { o.min_score_fraction = 0.8; o.match_score = 1; o.mismatch_penalty = 0; o.gap_open_penalty = 25; o.gap_extend_penalty = 25; }
  }
  if (o.quasi_coverage > 0 && !o.sensitive) o.sensitive = 1;
  if (index.empty() || r1.empty() || r2.empty()) { usage(); return 1; }

  rapmap_cuda_index_t* idx = nullptr;
  if (rapmap_cuda_index_load(index.c_str(), device, &idx) != RAPMAP_OK) die("loading index");
  FILE* out = nullptr;
  if (!noOutput) {
    out = outname.empty() ? stdout : std::fopen(outname.c_str(), "w");
    if (!out) { std::fprintf(stderr, "cannot open %s\n", outname.c_str()); return 1; }
    char* hdr = nullptr;
    uint64_t hl = 0;
    if (rapmap_cuda_sam_header(idx, &hdr, &hl) != RAPMAP_OK) die("SAM header");
    std::fwrite(hdr, 1, hl, out);
    rapmap_cuda_free(hdr);
  }
  Fastx f1, f2;
  if (!f1.open(r1) || !f2.open(r2)) { std::fprintf(stderr, "cannot open read files\n"); return 1; }

  rapmap_cuda_mapper_t* mapper = nullptr;
  uint32_t mapperMaxLen = 0;
  std::vector<uint8_t> s1, s2;
  std::vector<uint64_t> o1, o2;
  std::string n1, n2, name, seq;
  std::vector<rapmap_hit_t> hits(batch * 8 + 1024);
  std::vector<uint64_t> offs(batch + 1);
  uint64_t totalReads = 0, totalHits = 0;
  auto t0 = std::chrono::steady_clock::now();
  bool more = true;
  while (more) {
    s1.clear(); s2.clear(); n1.clear(); n2.clear();
    o1.assign(1, 0); o2.assign(1, 0);
    uint32_t maxLen = 0;
    uint64_t n = 0;
    while (n < batch) {
      if (!f1.next(name, seq)) { more = false; break; }
      n1 += name; n1 += '\0';
      s1.insert(s1.end(), seq.begin(), seq.end());
      o1.push_back(s1.size());
      if (seq.size() > maxLen) maxLen = static_cast<uint32_t>(seq.size());
      if (!f2.next(name, seq)) { std::fprintf(stderr, "mate files differ in length\n"); return 1; }
      n2 += name; n2 += '\0';
      s2.insert(s2.end(), seq.begin(), seq.end());
      o2.push_back(s2.size());
      if (seq.size() > maxLen) maxLen = static_cast<uint32_t>(seq.size());
      ++n;
    }
    if (n == 0) break;
    if (s1.empty()) s1.push_back(0);
    if (s2.empty()) s2.push_back(0);
    uint32_t need = maxLen < 31 ? 31 : maxLen;
    if (!mapper || need > mapperMaxLen) {
      if (mapper) rapmap_cuda_mapper_free(mapper);
      mapperMaxLen = (need + 15) / 16 * 16;
      if (rapmap_cuda_mapper_create(idx, &o, batch, mapperMaxLen, &mapper) != RAPMAP_OK) die("creating mapper");
    }
    rapmap_read_batch_t rb{};
    rb.seq1 = s1.data(); rb.off1 = o1.data(); rb.seq2 = s2.data(); rb.off2 = o2.data(); rb.n = n; rb.fixed_len = 0; rb.location = RAPMAP_LOC_HOST;
    rapmap_hit_batch_t hb{};
    hb.hits = hits.data(); hb.hits_capacity = hits.size(); hb.pair_offsets = offs.data(); hb.location = RAPMAP_LOC_HOST;
    int rc = rapmap_cuda_map_batch(mapper, &rb, &hb);
    if (rc == RAPMAP_ERR_CAPACITY) {
      hits.resize(hb.num_hits + 1024);
      hb.hits = hits.data(); hb.hits_capacity = hits.size();
      rc = rapmap_cuda_map_batch(mapper, &rb, &hb);
    }
    if (rc != RAPMAP_OK) die("map_batch");
    totalReads += hb.counters[0];
    totalHits += hb.counters[3];
    if (out) {
      char* sam = nullptr;
      uint64_t sl = 0;
      if (rapmap_cuda_format_sam(idx, &o, &rb, n1.c_str(), n2.c_str(), &hb, &sam, &sl) != RAPMAP_OK) die("format_sam");
      std::fwrite(sam, 1, sl, out);
      rapmap_cuda_free(sam);
    }
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (out && out != stdout) std::fclose(out);
  if (!quiet) {
    std::fprintf(stderr, "Elapsed time: %gs\n", secs);
    std::fprintf(stderr, "In total saw %llu reads.\nFinal # hits per read = %g\n", static_cast<unsigned long long>(totalReads),
                 totalReads ? static_cast<float>(totalHits) / static_cast<float>(totalReads) : 0.f);
  }
  if (mapper) rapmap_cuda_mapper_free(mapper);
  rapmap_cuda_index_free(idx);
  return 0;
}
