// `rapmap_b200 quasimap` — host front end keeping the reference CLI (flag table: reference src/RapMapSAMapper.cpp:992-1023)
// on top of the C-ABI.  The reference runs one FASTQ producer and -t consumers that each parse-map-format a 10,000-read chunk
// (src/RapMapSAMapper.cpp:801-909, src/FastxParser.cpp:229-328); here the stages are a pipeline around ONE mapper:
//
//   parser threads (one per mate file: gz inflate / line splitting into pinned chunk buffers)
//     -> main thread: rapmap_cuda_map_batch_async / rapmap_cuda_mapper_wait, up to rapmap_cuda_max_in_flight() chunks in flight (copy-in, kernels, copy-out overlap)
//     -> formatter: rapmap_cuda_format_sam_mt over -t host threads, written in input order
//
// Output is byte-identical to `rapmap quasimap -t 1` (paired reads, and unmated reads with -r).
#include <zlib.h>

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../include/rapmap_cuda.h"

namespace {

struct Fastx {
  gzFile f{nullptr};
  std::vector<char> buf;
  size_t pos{0}, len{0};
  bool open(const std::string& p) {
    f = gzopen(p.c_str(), "rb");
    if (f) gzbuffer(f, 1 << 20);
    buf.resize(4 << 20);
    return f != nullptr;
  }
  // next line as [b, e) inside the buffer when it does not straddle a refill, else assembled in `spill`
  bool getline(const char*& b, const char*& e, std::string& spill) {
    spill.clear();
    bool spilled = false;
    while (true) {
      if (pos == len) {
        int n = gzread(f, buf.data(), static_cast<unsigned>(buf.size()));
        if (n <= 0) {
          if (!spilled) return false;
          b = spill.data(); e = b + spill.size();
          return true;
        }
        len = static_cast<size_t>(n);
        pos = 0;
      }
      const char* s = buf.data() + pos;
      const char* nl = static_cast<const char*>(memchr(s, '\n', len - pos));
      if (nl) {
        if (spilled) { spill.append(s, nl - s); b = spill.data(); e = b + spill.size(); }
        else { b = s; e = nl; }
        pos += static_cast<size_t>(nl - s) + 1;
        if (e > b && e[-1] == '\r') --e;
        return true;
      }
      spill.append(s, len - pos);
      spilled = true;
      pos = len;
    }
  }
  ~Fastx() { if (f) gzclose(f); }
};

// One mate file of a chunk.
struct MateBuf {
  uint8_t* seq{nullptr};   // pinned
  uint64_t cap{0};
  std::vector<uint64_t> off;
  std::string names;       // '\0'-separated, kseq semantics: header up to the first whitespace
  std::vector<uint64_t> nameOff;  // start of every name in `names`
  uint64_t n{0};
  uint32_t maxLen{0};
  bool eof{false};
};

struct Chunk {
  MateBuf m[2];
  rapmap_hit_t* hits{nullptr};  // pinned
  uint64_t hitsCap{0};
  uint64_t* offs{nullptr};      // pinned
  rapmap_read_batch_t rb{};
  rapmap_hit_batch_t hb{};
  uint64_t n{0};
  bool last{false};
};

template <class T>
class Queue {
 public:
  void push(T v) { { std::lock_guard<std::mutex> g(mu_); q_.push_back(v); } cv_.notify_one(); }
  T pop() { std::unique_lock<std::mutex> g(mu_); cv_.wait(g, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
};

void die(const char* what) {
  std::fprintf(stderr, "[rapmap_b200] %s: %s\n", what, rapmap_cuda_last_error());
  std::exit(1);
}

void growSeq(MateBuf& mb, uint64_t need) {
  if (need <= mb.cap) return;
  uint64_t nc = mb.cap ? mb.cap : (1 << 20);
  while (nc < need) nc *= 2;
  auto* p = static_cast<uint8_t*>(rapmap_cuda_host_alloc(nc));
  if (!p) die("pinned read buffer");
  if (mb.seq) { std::memcpy(p, mb.seq, mb.off.empty() ? 0 : mb.off.back()); rapmap_cuda_host_free(mb.seq); }
  mb.seq = p; mb.cap = nc;
}

// Fills one mate of a chunk with up to `batch` records.
void parseMate(Fastx& fx, MateBuf& mb, uint64_t batch) {
  mb.off.assign(1, 0); mb.names.clear(); mb.nameOff.clear(); mb.n = 0; mb.maxLen = 0; mb.eof = false;
  std::string spill, spill2;
  const char *b, *e;
  while (mb.n < batch) {
    do {
      if (!fx.getline(b, e, spill)) { mb.eof = true; return; }
    } while (b == e);
    const bool fq = *b == '@';
    const char* nb = b + 1;
    const char* ne = nb;
    while (ne < e && *ne != ' ' && *ne != '\t') ++ne;
    mb.nameOff.push_back(mb.names.size());
    mb.names.append(nb, ne - nb);
    mb.names.push_back('\0');
    if (!fx.getline(b, e, spill)) { mb.eof = true; mb.names.resize(mb.nameOff.back()); mb.nameOff.pop_back(); return; }
    const uint64_t at = mb.off.back(), l = static_cast<uint64_t>(e - b);
    growSeq(mb, at + l + 16);
    std::memcpy(mb.seq + at, b, l);
    mb.off.push_back(at + l);
    if (l > mb.maxLen) mb.maxLen = static_cast<uint32_t>(l);
    ++mb.n;
    if (fq) { fx.getline(b, e, spill2); fx.getline(b, e, spill2); }
  }
}

void usage() {
  std::fprintf(stderr,
               "rapmap_b200 quasimap -i <index> (-1 <mates1> -2 <mates2> | -r <unmated>) [-o out.sam] [-t threads] [-s] [-m maxNumHits] [-z cov] [-f] [-n]\n"
               "        [--noOrphans] [--noDovetail] [--hardFilter] [--go N --ge N --mm N --ma N] [--dpBandwidth N] [--minScoreFrac F]\n"
               "        [--consensusSlack F] [--maxMMPExtension N] [--mimicBT2 | --mimicStrictBT2] [--noSensitive] [--noStrictCheck]\n"
               "        [--recoverOrphans] [--device D] [--batch PAIRS]\n");
}

} // namespace

int main(int argc, char** argv) {
  if (argc < 2 || std::strcmp(argv[1], "quasimap") != 0) {
    usage();
    return 1;
  }
  rapmap_cuda_opts_t o;
  rapmap_cuda_opts_default(&o);
  std::string index, r1, r2, ru, outname;
  bool noOutput = false, bt2 = false, strictBt2 = false, quiet = false;
  int device = 0;
  unsigned threads = 1;
  uint64_t batch = 1 << 18;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> std::string { if (i + 1 >= argc) { usage(); std::exit(1); } return argv[++i]; };
    if (a == "-i" || a == "--index") index = val();
    else if (a == "-1" || a == "--leftMates") r1 = val();
    else if (a == "-2" || a == "--rightMates") r2 = val();
    else if (a == "-r" || a == "--unmatedReads") ru = val();
    else if (a == "-o" || a == "--output") outname = val();
    else if (a == "-t" || a == "--numThreads") threads = static_cast<unsigned>(std::max(1, std::stoi(val())));  // host threads of the SAM formatter
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
  const bool paired = ru.empty();
  if (index.empty() || (paired && (r1.empty() || r2.empty())) || batch == 0) { usage(); return 1; }
  if (!paired && (!r1.empty() || !r2.empty())) { std::fprintf(stderr, "give either -1/-2 or -r\n"); return 1; }  // src/RapMapSAMapper.cpp:1081-1107

  rapmap_cuda_index_t* idx = nullptr;
  if (rapmap_cuda_index_load(index.c_str(), device, &idx) != RAPMAP_OK) die("loading index");
  FILE* out = nullptr;
  if (!noOutput) {
    out = outname.empty() ? stdout : std::fopen(outname.c_str(), "w");
    if (!out) { std::fprintf(stderr, "cannot open %s\n", outname.c_str()); return 1; }
    std::setvbuf(out, nullptr, _IOFBF, 8 << 20);
    char* hdr = nullptr;
    uint64_t hl = 0;
    if (rapmap_cuda_sam_header(idx, &hdr, &hl) != RAPMAP_OK) die("SAM header");
    std::fwrite(hdr, 1, hl, out);
    rapmap_cuda_free(hdr);
  }
  Fastx fx[2];
  if (!fx[0].open(paired ? r1 : ru) || (paired && !fx[1].open(r2))) { std::fprintf(stderr, "cannot open read files\n"); return 1; }

  // ---- chunk pool and queues
  constexpr int kChunks = 6;
  std::vector<Chunk> pool(kChunks);
  Queue<Chunk*> freeQ, parsedQ, mappedQ;
  for (auto& c : pool) {
    c.hitsCap = batch * 8 + 1024;
    c.hits = static_cast<rapmap_hit_t*>(rapmap_cuda_host_alloc(c.hitsCap * sizeof(rapmap_hit_t)));
    c.offs = static_cast<uint64_t*>(rapmap_cuda_host_alloc((batch + 1) * 8));
    if (!c.hits || !c.offs) die("pinned result buffers");
    freeQ.push(&c);
  }
  auto t0 = std::chrono::steady_clock::now();

  // ---- parser: the two mate files are inflated / split by two threads at the same time
  std::thread parser([&] {
    for (;;) {
      Chunk* c = freeQ.pop();
      std::thread second;
      if (paired) second = std::thread([&] { parseMate(fx[1], c->m[1], batch); });
      parseMate(fx[0], c->m[0], batch);
      if (paired) {
        second.join();
        if (c->m[0].n != c->m[1].n) { std::fprintf(stderr, "mate files differ in length\n"); std::exit(1); }
      }
      c->n = c->m[0].n;
      c->last = c->m[0].eof || c->n < batch;
      parsedQ.push(c);
      if (c->last) return;
    }
  });

  // ---- formatter / writer (input order: chunks arrive in the order they were mapped)
  uint64_t totalReads = 0, totalHits = 0;
  std::thread formatter([&] {
    for (;;) {
      Chunk* c = mappedQ.pop();
      if (!c) return;
      if (out && c->n > 0) {
        // -t threads format disjoint ranges of the chunk (views into the same buffers); the pieces are written in order
        const unsigned T = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(threads, (c->n + 2047) / 2048)));
        std::vector<char*> piece(T, nullptr);
        std::vector<uint64_t> plen(T, 0);
        std::vector<int> rcs(T, 0);
        auto work = [&](unsigned t) {
          const uint64_t a = c->n * t / T, b = c->n * (t + 1) / T;
          rapmap_read_batch_t rb = c->rb;
          rb.off1 += a; rb.n = b - a;
          if (paired) rb.off2 += a;
          rapmap_hit_batch_t hb = c->hb;
          hb.pair_offsets += a;
          rcs[t] = rapmap_cuda_format_sam(idx, &o, &rb, c->m[0].names.data() + c->m[0].nameOff[a], paired ? c->m[1].names.data() + c->m[1].nameOff[a] : nullptr, &hb,
                                          &piece[t], &plen[t]);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        for (unsigned t = 0; t < T; ++t) {
          if (rcs[t] != RAPMAP_OK) die("format_sam");
          std::fwrite(piece[t], 1, plen[t], out);
          rapmap_cuda_free(piece[t]);
        }
      }
      const bool last = c->last;
      freeQ.push(c);
      if (last) return;
    }
  });

  // ---- mapping: several chunks in flight on one mapper
  rapmap_cuda_mapper_t* mapper = nullptr;
  uint32_t mapperMaxLen = 0;
  std::deque<Chunk*> inFlight;
  auto account = [&](Chunk* c) { totalReads += c->hb.counters[0]; totalHits += c->hb.counters[3]; mappedQ.push(c); };
  auto redo = [&](Chunk* c) {  // more hits than the chunk's buffer holds: grow it and map the chunk again (nothing else is in flight)
    rapmap_cuda_host_free(c->hits);
    c->hitsCap = c->hb.num_hits + 1024;
    c->hits = static_cast<rapmap_hit_t*>(rapmap_cuda_host_alloc(c->hitsCap * sizeof(rapmap_hit_t)));
    if (!c->hits) die("pinned result buffers");
    c->hb.hits = c->hits; c->hb.hits_capacity = c->hitsCap;
    if (rapmap_cuda_map_batch(mapper, &c->rb, &c->hb) != RAPMAP_OK) die("map_batch");
  };
  auto collect = [&]() {
    Chunk* c = inFlight.front();
    inFlight.pop_front();
    int rc = rapmap_cuda_mapper_wait(mapper);
    if (rc == RAPMAP_OK) { account(c); return; }
    if (rc != RAPMAP_ERR_CAPACITY) die("map_batch");
    // drain the chunk behind it first (its result waits in its own buffers), then repeat, and forward both in input order
    std::vector<std::pair<Chunk*, int>> rest;
    while (!inFlight.empty()) { Chunk* d = inFlight.front(); inFlight.pop_front(); rest.emplace_back(d, rapmap_cuda_mapper_wait(mapper)); }
    redo(c);
    account(c);
    for (auto& dr : rest) {
      if (dr.second == RAPMAP_ERR_CAPACITY) redo(dr.first);
      else if (dr.second != RAPMAP_OK) die("map_batch");
      account(dr.first);
    }
  };
  for (;;) {
    Chunk* c = parsedQ.pop();
    uint32_t need = std::max<uint32_t>(31, std::max(c->m[0].maxLen, c->m[1].maxLen));
    if (!mapper || need > mapperMaxLen) {
      while (!inFlight.empty()) collect();
      if (mapper) rapmap_cuda_mapper_free(mapper);
      mapperMaxLen = std::min<uint32_t>(1000, (need + 31) / 32 * 32);
      if (rapmap_cuda_mapper_create(idx, &o, batch, mapperMaxLen, &mapper) != RAPMAP_OK) die("creating mapper");
    }
    if (c->n > 0) {
      for (int k = 0; k < (paired ? 2 : 1); ++k) growSeq(c->m[k], 16);
      c->rb = rapmap_read_batch_t{c->m[0].seq, c->m[0].off.data(), paired ? c->m[1].seq : nullptr, paired ? c->m[1].off.data() : nullptr, c->n, 0, RAPMAP_LOC_HOST};
      c->hb = rapmap_hit_batch_t{c->hits, c->hitsCap, c->offs, 0, {0, 0, 0, 0, 0}, RAPMAP_LOC_HOST};
      if (inFlight.size() == rapmap_cuda_max_in_flight()) collect();
      if (rapmap_cuda_map_batch_async(mapper, &c->rb, &c->hb) != RAPMAP_OK) die("map_batch_async");
      inFlight.push_back(c);
    } else {
      while (!inFlight.empty()) collect();
      mappedQ.push(c);
    }
    if (c->last) break;
  }
  while (!inFlight.empty()) collect();
  parser.join();
  formatter.join();
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (out && out != stdout) std::fclose(out);
  else if (out) std::fflush(out);
  if (!quiet) {
    std::fprintf(stderr, "Elapsed time: %gs\n", secs);
    std::fprintf(stderr, "In total saw %llu reads.\nFinal # hits per read = %g\n", static_cast<unsigned long long>(totalReads),
                 totalReads ? static_cast<float>(totalHits) / static_cast<float>(totalReads) : 0.f);
  }
  if (mapper) rapmap_cuda_mapper_free(mapper);
  for (auto& c : pool) { rapmap_cuda_host_free(c.hits); rapmap_cuda_host_free(c.offs); for (auto& mb : c.m) rapmap_cuda_host_free(mb.seq); }
  rapmap_cuda_index_free(idx);
  return 0;
}
