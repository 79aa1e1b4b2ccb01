// TEST INFRASTRUCTURE.  The stub of INTEGRATION.md §1, compiled against the UNMODIFIED reference headers and linked with the
// reference's own objects (oracle/build_ref.sh -> oracle/_ref/adapter_sam): the reference's FASTQ parser feeds ReadGroup
// chunks to rapmap_b200::BatchMapper (include/rapmap_b200/adapter.hpp), and the std::vector<QuasiAlignment> it returns go
// through the reference's own writeSAMHeader / writeAlignmentsToStream / writeUnalignedPairToStream
// (src/RapMapUtils.cpp:137-588, include/RapMapUtils.hpp:95-131).  tests/test_gpu_adapter.py md5-checks the output against
// the golden SAM: that is what "drops into the reference's mapping path" means at the source level.
//
// usage: adapter_sam <index dir> <out.sam> [-s] (-1 reads_1.fastq -2 reads_2.fastq | -r reads.fastq) [--chunk N]
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "FastxParser.hpp"
#include "PairAlignmentFormatter.hpp"
#include "RapMapSAIndex.hpp"
#include "RapMapUtils.hpp"
#include "SingleAlignmentFormatter.hpp"
#include "rapmap_b200/adapter.hpp"
#include "spdlog/sinks/ostream_sink.h"
#include "spdlog/sinks/stdout_sinks.h"
#include "spdlog/spdlog.h"

using SAIndex32BitDense = RapMapSAIndex<int32_t, RegHashT<uint64_t, rapmap::utils::SAInterval<int32_t>, rapmap::utils::KmerKeyHasher>>;

int main(int argc, char** argv) {
  if (argc < 5) { std::cerr << "usage: adapter_sam <index> <out.sam> [-s] (-1 a.fq -2 b.fq | -r u.fq) [--chunk N]\n"; return 2; }
  const std::string indexDir = argv[1], outName = argv[2];
  std::string r1, r2, ru;
  bool selAln = false;
  size_t chunk = 10000;  // the reference's miniBatchSize, src/RapMapSAMapper.cpp:853
  for (int i = 3; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-s") selAln = true;
    else if (a == "-1" && i + 1 < argc) r1 = argv[++i];
    else if (a == "-2" && i + 1 < argc) r2 = argv[++i];
    else if (a == "-r" && i + 1 < argc) ru = argv[++i];
    else if (a == "--chunk" && i + 1 < argc) chunk = std::stoul(argv[++i]);
    else { std::cerr << "unknown argument " << a << "\n"; return 2; }
  }
  // the reference's loaders log through a registered spdlog logger (src/RapMapSAMapper.cpp:1178-1180)
  auto consoleSink = std::make_shared<spdlog::sinks::stderr_sink_mt>();
  auto consoleLog = spdlog::create("stderrLog", {consoleSink});
  // the reference's index object: transcript names / lengths for its SAM writers
  SAIndex32BitDense rmi;
  if (!rmi.load(indexDir)) { std::cerr << "reference index load failed\n"; return 1; }
  rapmap_cuda_index_t* gidx = nullptr;
  if (rapmap_cuda_index_load(indexDir.c_str(), 0, &gidx) != RAPMAP_OK) { std::cerr << rapmap_cuda_last_error() << "\n"; return 1; }
  rapmap_cuda_opts_t o;
  if (selAln) rapmap_cuda_opts_selaln(&o); else rapmap_cuda_opts_default(&o);

  // output exactly as mapReads sets it up (src/RapMapSAMapper.cpp:842-846): a "%v" spdlog logger over the file stream; the
  // logger overload of writeSAMHeader prints rapmap::version (the std::ostream overload carries a stale literal)
  std::ofstream out(outName, std::ios::binary);
  auto outputSink = std::make_shared<spdlog::sinks::ostream_sink_mt>(out);
  auto outLog = std::make_shared<spdlog::logger>("rapmap::outLog", outputSink);
  outLog->set_pattern("%v");
  rapmap::utils::writeSAMHeader(rmi, outLog);
  auto flushChunk = [&](fmt::MemoryWriter& ss) {  // src/RapMapSAMapper.cpp:350-363: drop the trailing newline, the logger adds one
    std::string outStr(ss.str());
    if (!outStr.empty()) { outStr.pop_back(); outLog->info(std::move(outStr)); }
    ss.clear();
  };
  rapmap::utils::HitCounters hctr;
  std::vector<std::vector<rapmap::utils::QuasiAlignment>> hits;
  fmt::MemoryWriter sstream;
  try {
    rapmap_b200::BatchMapper mapper(gidx, o, chunk, 1000);
    if (!r1.empty()) {
      PairAlignmentFormatter<SAIndex32BitDense*> formatter(&rmi);
      fastx_parser::FastxParser<fastx_parser::ReadPair> parser({r1}, {r2}, 1, 1, chunk);
      parser.start();
      auto rg = parser.getReadGroup();
      while (parser.refill(rg)) {
        mapper.mapPairs(rg, hits, hctr);
        size_t i = 0;
        for (auto& rp : rg) {
          auto& jointHits = hits[i++];
          if (!jointHits.empty()) rapmap::utils::writeAlignmentsToStream(rp, formatter, hctr, jointHits, sstream);
          else rapmap::utils::writeUnalignedPairToStream(rp, sstream);
        }
        flushChunk(sstream);
      }
      parser.stop();
    } else {
      SingleAlignmentFormatter<SAIndex32BitDense*> formatter(&rmi);
      fastx_parser::FastxParser<fastx_parser::ReadSeq> parser({ru}, 1, 1, chunk);
      parser.start();
      auto rg = parser.getReadGroup();
      while (parser.refill(rg)) {
        mapper.mapSingles(rg, hits, hctr);
        size_t i = 0;
        for (auto& r : rg) {
          auto& h = hits[i++];
          if (!h.empty()) rapmap::utils::writeAlignmentsToStream(r, formatter, hctr, h, sstream);
          else rapmap::utils::writeUnalignedSingleToStream(r, sstream);
        }
        flushChunk(sstream);
      }
      parser.stop();
    }
  } catch (const std::exception& e) {
    std::cerr << "adapter_sam: " << e.what() << "\n";
    return 1;
  }
  outLog->flush();
  out.close();
  std::cerr << "adapter_sam: reads " << hctr.numReads << " totHits " << hctr.totHits << "\n";
  rapmap_cuda_index_free(gidx);
  return 0;
}
