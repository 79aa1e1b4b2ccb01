// C-ABI of the B200 quasi-mapping engine (include/rapmap_cuda.h): index image construction and upload,
// per-stream mapper state, and the kernel pipeline behind rapmap_cuda_map_batch.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rapmap_cuda.h"
#include "hits_to_mappings.cuh"
#include "index_loader.hpp"
#include "kernels.cuh"
#include "merge_pairs.cuh"
#include "sa_collect_lane.cuh"
#include "sam_writer.hpp"
#include "sel_aln.cuh"

using namespace rapmap_b200;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(call)                                                                                     \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(RAPMAP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                \
  } while (0)

// No exception crosses the C-ABI: allocation failures and the like become error codes.
template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail(RAPMAP_ERR_IO, "out of host memory (or a corrupt size field in the index)");
  } catch (const std::exception& e) {
    return fail(RAPMAP_ERR_IO, std::string("exception: ") + e.what());
  } catch (...) {
    return fail(RAPMAP_ERR_IO, "unknown exception");
  }
}

inline uint64_t align256(uint64_t x) { return (x + 255) / 256 * 256; }

struct CastU64 {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return v; }
};

// ---- hash table construction on the device ------------------------------------------------------
__global__ void build_table_kernel(const KmerRecord* recs, uint64_t n, uint4* table, uint64_t mask) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    KmerRecord r = recs[i];
    uint64_t s = mix64(r.kmer) & mask & ~1ULL;
    while (true) {
      unsigned long long* keyp = reinterpret_cast<unsigned long long*>(table + s);
      unsigned long long old = atomicCAS(keyp, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(r.kmer));
      if (old == kEmptyKey || old == r.kmer) {
        reinterpret_cast<int32_t*>(table + s)[2] = r.begin;
        reinterpret_cast<int32_t*>(table + s)[3] = r.end;
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

// -p index: the dense table derived from the perfect hash.  FrugalBooMap::find(q) can only succeed when q is the k-mer the
// text holds at SA[data_[lookup(q)]], i.e. when q is one of the keys key_i = textWord(SA[data_[i]]), i < |data_|.  So the
// table gets (key_i -> find(key_i)) for every i whose find succeeds - found or not is decided by the SAME device function the
// lookups used before - and answers every later query exactly as the walk would: a query that hits in the walk equals its
// own key_i and was inserted; a query in the table hit in the walk by construction.
__global__ void derive_table_from_phf_kernel(DeviceIndex ix, uint4* table, uint64_t mask) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < ix.phfNumData; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const int32_t ind = ix.phfData[i];
    const uint64_t key = phfTextWord(ix, ix.SA[ind]);
    const int2 r = phfFindImpl(ix, key);
    if (r.x < 0) continue;
    uint64_t s = mix64(key) & mask & ~1ULL;
    while (true) {
      unsigned long long* keyp = reinterpret_cast<unsigned long long*>(table + s);
      const unsigned long long old = atomicCAS(keyp, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
      if (old == kEmptyKey || old == key) {
        reinterpret_cast<int32_t*>(table + s)[2] = r.x;
        reinterpret_cast<int32_t*>(table + s)[3] = r.y;
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

// SA entry -> {transcript id, position in the transcript} (RapMapSAIndex::transcriptAtPosition + txpOffsets, once per entry)
__global__ void build_sa_tidpos_kernel(DeviceIndex ix, uint2* out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < ix.n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t g = ix.SA[i];
    const uint32_t tid = transcriptAt(ix, g);
    out[i] = make_uint2(tid, static_cast<uint32_t>(g - ix.txpOffsets[tid]));
  }
}

__global__ void build_filter_kernel(const KmerRecord* recs, uint64_t n, uint32_t k, uint32_t* filter, uint32_t shift) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    uint64_t word;
    uint32_t mask;
    filterInsertBits(recs[i].kmer, k, shift, word, mask);
    atomicOr(filter + word, mask);
  }
}

__global__ void build_filter_from_text_kernel(const uint8_t* text, uint64_t n, uint32_t k, uint32_t* filter, uint32_t shift) {
  for (uint64_t p = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p + k <= n; p += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    uint64_t w = 0;
    bool ok = true;
    for (uint32_t j = 0; j < k; ++j) {  // same encoding as the verification step of phfFindImpl
      const uint8_t ch = text[p + j];
      uint32_t cd;
      if (ch == 'A') cd = 0; else if (ch == 'C') cd = 1; else if (ch == 'G') cd = 2; else if (ch == 'T') cd = 3; else { ok = false; break; }
      w |= static_cast<uint64_t>(cd) << (2 * (k - 1 - j));
    }
    if (!ok) continue;
    uint64_t word;
    uint32_t mask;
    filterInsertBits(w, k, shift, word, mask);
    atomicOr(filter + word, mask);
  }
}

// Packed-text records from the ASCII text (device_index.cuh: TextRec).  *bad is raised when a character lies outside
// '$'..'z': the packed compare orders the search sentinels '#' and '{' against every text character without looking.
__global__ void build_text2_kernel(const uint8_t* text, uint64_t n, TextRec* recs, uint64_t numRecs, uint32_t* bad) {
  for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < numRecs; j += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    uint64_t c[2] = {0, 0};
    uint32_t inv[2] = {0, 0};
    for (int h = 0; h < 2; ++h) {
      for (int b = 0; b < 32; ++b) {
        const uint64_t p = j * 32 + static_cast<uint64_t>(h) * 32 + b;
        uint32_t code = 0;
        bool ok = false;
        if (p < n) {
          const uint8_t ch = text[p];
          if (ch == 'A') { code = 0; ok = true; } else if (ch == 'C') { code = 1; ok = true; }
          else if (ch == 'G') { code = 2; ok = true; } else if (ch == 'T') { code = 3; ok = true; }
          else if (ch < '$' || ch > 'z') *bad = 1u;
        }
        c[h] |= static_cast<uint64_t>(code) << (62 - 2 * b);
        if (!ok) inv[h] |= 1u << b;
      }
    }
    TextRec r;
    r.c0lo = static_cast<uint32_t>(c[0]); r.c0hi = static_cast<uint32_t>(c[0] >> 32);
    r.c1lo = static_cast<uint32_t>(c[1]); r.c1hi = static_cast<uint32_t>(c[1] >> 32);
    r.inv0 = inv[0]; r.inv1 = inv[1]; r.pad0 = 0; r.pad1 = 0;
    recs[j] = r;
  }
}

// Result copy-out without a host round trip: the number of records is only known on the device, so a kernel moves them -
// to the caller's device buffer, or straight into its pinned host buffer over PCIe (16-byte stores, coalesced).  Nothing is
// written when the records do not fit `cap` (the host reports RAPMAP_ERR_CAPACITY with the count).
// `skip` (a multiple of 4 records): the leading records a copy engine has already been asked to move (pinned host buffers,
// see enqueueAttempt); offSrc == nullptr: the offsets went the same way.
__global__ void __launch_bounds__(256) copy_out_kernel(const rapmap_hit_t* __restrict__ src, rapmap_hit_t* __restrict__ dst, uint64_t cap,
                                                       const uint64_t* __restrict__ totalPtr, const uint64_t* __restrict__ offSrc, uint64_t* __restrict__ offDst,
                                                       uint64_t nOff, uint64_t skip) {
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x, stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  if (offSrc != nullptr)
    for (uint64_t i = tid; i < nOff; i += stride) offDst[i] = offSrc[i];
  const uint64_t total = *totalPtr;
  if (total > cap || dst == nullptr || total <= skip) return;
  const uint64_t words = total * (sizeof(rapmap_hit_t) / 4), w0 = skip * (sizeof(rapmap_hit_t) / 4);  // w0 is a multiple of 4 words
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const uint64_t n4 = words / 4;
    for (uint64_t i = w0 / 4 + tid; i < n4; i += stride) d4[i] = s4[i];
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
    for (uint64_t i = n4 * 4 + tid; i < words; i += stride) d1[i] = s1[i];
  } else {
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
    for (uint64_t i = w0 + tid; i < words; i += stride) d1[i] = s1[i];
  }
}

// Device-visible alias of a caller buffer: device memory as it is, pinned / registered host memory through its mapped
// address; nullptr for pageable host memory (then the copy needs cudaMemcpyAsync and a host round trip for the size).
void* deviceAlias(const void* p, int location) {
  if (!p) return nullptr;
  if (location == RAPMAP_LOC_DEVICE) return const_cast<void*>(p);
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type == cudaMemoryTypeHost && a.devicePointer) return a.devicePointer;
  return nullptr;
}

DeviceIndex viewOf(const uint8_t* blob, const ImageHeader& h) {
  DeviceIndex d;
  d.SA = reinterpret_cast<const int32_t*>(blob + h.offSA);
  d.text = blob + h.offText;
  d.text2 = h.offText2 ? reinterpret_cast<const TextRec*>(blob + h.offText2) : nullptr;
  d.filter = h.offFilter ? reinterpret_cast<const uint32_t*>(blob + h.offFilter) : nullptr;
  d.filterShift = 64;
  for (uint64_t w = h.filterWords; w > 1; w >>= 1) --d.filterShift;
  d.rank = reinterpret_cast<const uint4*>(blob + h.offRank);
  d.txpOffsets = reinterpret_cast<const int32_t*>(blob + h.offTxpOffsets);
  d.txpLens = reinterpret_cast<const int32_t*>(blob + h.offTxpLens);
  d.saTidPos = reinterpret_cast<const uint2*>(blob + h.offSaTidPos);
  d.table = reinterpret_cast<const uint4*>(blob + h.offTable);
  d.tableMask = h.tableSlots - 1;
  d.n = static_cast<int64_t>(h.n);
  d.k = h.k;
  d.numTxp = static_cast<uint32_t>(h.numTxp);
  d.hashKind = (h.hashKind && !h.phfTableDerived) ? 1u : 0u;  // what the kernels walk: the BooPHF arrays only while no table answers for them
  d.phfLevels = h.phfLevels; d.phfNumFinal = static_cast<uint32_t>(h.phfNumFinal);
  d.phfNumOverflow = static_cast<uint32_t>(h.phfNumOverflow); d.phfLastRank = h.phfLastRank; d.phfNumData = h.phfNumData;
  d.phfLv = reinterpret_cast<const PhfLevelDev*>(blob + h.offPhfLevels);
  d.phfBits = reinterpret_cast<const uint64_t*>(blob + h.offPhfBits);
  d.phfRanks = reinterpret_cast<const uint64_t*>(blob + h.offPhfRanks);
  d.phfFinal = reinterpret_cast<const ulonglong2*>(blob + h.offPhfFinal);
  d.phfData = reinterpret_cast<const int32_t*>(blob + h.offPhfData);
  d.phfLens = blob + h.offPhfLens;
  d.phfOverflow = reinterpret_cast<const int2*>(blob + h.offPhfOverflow);
  return d;
}

} // namespace

struct rapmap_cuda_index {
  int device{0};
  uint8_t* blob{nullptr};
  bool ownsBlob{true};
  ImageHeader hdr{};
  DeviceIndex view{};
  std::vector<std::string> names;
  std::vector<int32_t> lens;
};

// Batches a mapper keeps in flight (kDepth slots): while batch c computes, batch c+1's reads come in over PCIe and batch
// c-1's results go out - on three streams of ONE mapper, so the kernels of consecutive batches never compete for the SMs (three mappers on
// three streams lost a quarter of the device-resident rate to that, profiles/r02d_e2e_diag.txt).
#ifndef RAPMAP_DEPTH
#define RAPMAP_DEPTH 3  // with 2 the copy-in of chunk c+2 could only start after the copy-out of chunk c: copy-out + copy-in ~ one chunk of compute, no slack
#endif
static constexpr int kDepth = RAPMAP_DEPTH;

// pinned block the compute stream writes at the end of an attempt; the host reads it after the batch's synchronisation
struct StageBlock { uint32_t ctl[4]; uint64_t mergeTotal; uint64_t selTotal; uint32_t dpJobs[4]; Counters5 counters; };

struct BatchSlot {
  // reads staged from the host
  uint8_t* dSeq[2]{nullptr, nullptr};
  uint64_t* dOff[2]{nullptr, nullptr};
  // results (per slot: batch c+1 computes while batch c goes out)
  uint64_t* dPairOff{nullptr};
  rapmap_hit_t* dHits{nullptr};
  rapmap_hit_t* dSelOut{nullptr};   // survivors of the selective-alignment filter
  StageBlock* hStage{nullptr};
  cudaEvent_t ev[12]{};             // stage boundaries (timing)
  cudaEvent_t evIn{nullptr}, evCompute{nullptr};
  cudaEvent_t evOut{nullptr};       // cudaEventBlockingSync: a waiting host thread sleeps instead of spinning
  BatchView view{};
  bool paired{false};
  bool inFlight{false};
  bool rerun{false};                // an arena was re-allocated under this batch: its attempt has to be repeated
  rapmap_hit_batch_t* out{nullptr};
  bool directOut{false};            // results leave by copy_out_kernel (device or pinned host buffers)
  void* outHitsDev{nullptr};
  void* outOffDev{nullptr};
  uint32_t launches{0};
};

struct rapmap_cuda_mapper {
  const rapmap_cuda_index* idx{nullptr};
  rapmap_cuda_opts_t opts{};
  DevOpts dopts{};
  uint64_t maxBatch{0};
  uint32_t maxReadLen{0};
  cudaStream_t stream{nullptr};      // compute
  cudaStream_t sIn{nullptr}, sOut{nullptr};
  int numSMs{0};
  uint64_t seqCap{0};
  BatchSlot slots[kDepth];
  uint64_t submitted{0}, collected{0};  // batch sequence numbers: slot = seq % kDepth
  // stage 1: SA lookup (lane per read)
  ReadSummary* dSumm{nullptr};
  IntervalRec* dIvArena{nullptr};
  uint32_t ivCap{0};
  uint64_t scratchSlotsK1{0};
  uint4* dPacked{nullptr};
  uint4* dKmask{nullptr};
  uint32_t* dOrder{nullptr};
  uint32_t* dClassCtl{nullptr};
  IntervalRec* dIvScratch{nullptr};
  uint32_t ivStride{0};
  uint32_t* dVoteScratch{nullptr};
  uint32_t voteWords{0}, laneWords{0}, laneSmem{0}, pmax{0};
  int gridLane{0};
  bool masksInGlobal{false};
  void (*laneKernel)(LaneParams){nullptr};   // sa_collect_lane_kernel instantiation for this index flavour / flag set
  // stage 2: hit resolution
  QASummary* dQSumm{nullptr};
  QARec* dQaArena{nullptr};
  uint32_t qaCap{0};
  int32_t* dPosPool{nullptr};
  uint32_t posCap{0};
  uint8_t* dScratch{nullptr};
  uint32_t scratchEntries{0};
  uint64_t scratchStride{0};
  uint32_t smemEntries{64};
  int gridMap{0};
  uint32_t mapSmem{0};
  bool laneMap{false};
  bool chainLaneMap{false};
  int gridLaneMap{0};
  int gridChain[2]{0, 0};
  uint32_t* dOrder2{nullptr};    // reads grouped by SA-entry count (hit resolution with chaining)
  uint32_t* dK2Class{nullptr};
  // stage 3: mate merge
  uint32_t* dPairCount{nullptr};
  uint64_t hitsCap{0};
  // stage 4: selective alignment
  SelAlnWork selaln{};
  SelAlnLaunch selLaunch{};
  void* dCubTemp{nullptr};
  size_t cubTempBytes{0};
  // control words: [0] interval cursor, [1] qa cursor, [2] pos cursor, [3] status, [4] read cursor of the lane kernel
  uint32_t* dCtl{nullptr};
  Counters5* dCounters{nullptr};
  rapmap_cuda_timing_t timing{};
  uint64_t lastReads{0};
  double hitsPerPair{0.0};          // records per pair of the last collected batch (sizes the copy-engine part of the next copy-outs)
};

static constexpr int kWarps = 8;
#ifndef RAPMAP_LANE_THREADS
#define RAPMAP_LANE_THREADS 256
#endif
#ifndef RAPMAP_LANE_MINB
#define RAPMAP_LANE_MINB 3
#endif
#ifndef RAPMAP_MAPLANE_CAP
#define RAPMAP_MAPLANE_CAP 16
#endif
// the size ranges of the chaining lane kernel: (threads per block, strip entries); ~68 KB of shared memory per block each.
// A third range (32 threads, 64 entries) ran at 3 warps per SM: 6.4 ms for the 22 % of reads with 33-64 entries, which the
// warp-per-read kernel resolves in ~2.5 ms (profiles/r02f_launches_selaln.csv)
#define RAPMAP_CHAIN_CLASSES(X) X(128, 16) X(64, 32)
static constexpr int kMapLaneThreads = 128;                // lane-per-read hit resolution
static constexpr int kMapLaneCap = RAPMAP_MAPLANE_CAP;     // SA entries per read it takes
static constexpr uint32_t kMapLaneSmem = 2u * kMapLaneCap * kMapLaneThreads * 8u;
static constexpr int kLaneThreads = RAPMAP_LANE_THREADS;  // lane-per-read SA-lookup kernel: threads per block
static constexpr int kLaneMinBlocks = RAPMAP_LANE_MINB;   // 3 x 256 threads, 80 registers: the 64-register build spills and is 13 % slower

extern "C" {

const char* rapmap_cuda_last_error(void) { return g_err.c_str(); }

void rapmap_cuda_opts_default(rapmap_cuda_opts_t* o) {
  std::memset(o, 0, sizeof(*o));
  o->max_num_hits = 200;
  o->quasi_coverage = 0.0;
  o->sensitive = 1;
  o->strict_check = 1;
  o->consensus_slack = 0.2f;
  o->min_score_fraction = 0.65;
  o->match_score = 2;
  o->mismatch_penalty = -4;
  o->gap_open_penalty = 4;
  o->gap_extend_penalty = 2;
  o->dp_bandwidth = 15;
  o->max_mmp_extension = 7;
}

void rapmap_cuda_opts_selaln(rapmap_cuda_opts_t* o) {
  rapmap_cuda_opts_default(o);
  o->sel_aln = 1;
}

// RAPMAP_B200_PHF=walk: -p indexes keep walking the BooPHF arrays on the device (no derived table; the smaller image).
static bool phfWalkOnly() {
  const char* t = std::getenv("RAPMAP_B200_PHF");
  return t && std::string(t) == "walk";
}

static int indexLoadImpl(const char* index_dir, int device, rapmap_cuda_index_t** out) {
  if (!index_dir || !out) return fail(RAPMAP_ERR_ARG, "null argument");
  *out = nullptr;
  HostIndex h;
  std::string err;
  if (!h.load(index_dir, err)) return fail(h.unsupported ? RAPMAP_ERR_UNSUPPORTED : RAPMAP_ERR_IO, err);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(RAPMAP_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this engine has no CPU path");
  if (device < 0 || device >= ndev) return fail(RAPMAP_ERR_ARG, "bad device ordinal");
  CU_TRY(cudaSetDevice(device));

  const uint64_t n = h.SA.size();
  const uint64_t T = h.txpOffsets.size();
  uint64_t slots = 64;
  while (slots < 2 * h.kmers.size()) slots <<= 1;
  // -p index: by default the lookups are served by a dense table DERIVED from the perfect hash once the image is up (180 GB of
  // HBM: 4.3 GB buy back the 3.6x that the BooPHF walk costs the SA-lookup kernel); RAPMAP_B200_PHF=walk keeps the walk
  const bool deriveTable = h.perfectHash && !phfWalkOnly();
  if (h.perfectHash) {
    slots = 16;
    if (deriveTable) while (slots < 2 * h.phf.data.size()) slots <<= 1;
  }
  const uint64_t rankWords = n / 64 + 1;

  ImageHeader hdr{};
  hdr.magic = kImageMagic;
  hdr.n = n; hdr.numTxp = T; hdr.tableSlots = slots; hdr.numKmers = h.kmers.size(); hdr.k = h.k;
  uint64_t off = align256(sizeof(ImageHeader));
  hdr.offSA = off; off = align256(off + n * 4);
  hdr.offText = off; off = align256(off + n + 256);
  hdr.offRank = off; off = align256(off + rankWords * 16);
  hdr.offTxpOffsets = off; off = align256(off + T * 4);
  hdr.offTxpLens = off; off = align256(off + T * 4);
  hdr.offTable = off; off = align256(off + slots * 16);
  const uint64_t numKeys = h.perfectHash ? h.phf.data.size() : h.kmers.size();
  if (numKeys > 0) {  // ~6 bits per k-mer, capped at 64 MB (L2-resident)
    uint64_t words = 1024;
    while (words * 32 < 6 * numKeys && words < (1ull << 24)) words <<= 1;
    hdr.filterWords = words;
    hdr.offFilter = off; off = align256(off + words * 4);
  }
  const uint64_t text2Recs = n / 32 + 2;
  hdr.offText2 = off; off = align256(off + text2Recs * sizeof(TextRec));
  // -p index: level table, concatenated bitsets and rank samples, _final_hash, data_, lens_, overflow_
  uint64_t phfWords = 0, phfRanks = 0;
  std::vector<PhfLevelDev> lvDev;
  if (h.perfectHash) {
    hdr.hashKind = 1;
    hdr.phfLevels = static_cast<uint32_t>(h.phf.nbLevels);
    hdr.phfLastRank = h.phf.lastBitsetRank; hdr.phfNumData = h.phf.data.size(); hdr.phfNumFinal = h.phf.finalHash.size();
    hdr.phfNumOverflow = h.phf.overflow.size();
    for (auto& lv : h.phf.levels) {
      PhfLevelDev d{lv.hashDomain, phfWords, phfRanks, 0};
      lvDev.push_back(d);
      phfWords += lv.bits.size() + 1;   // one spare word: rank() may read bits[word_idx] for word_idx == nchar - 1
      phfRanks += lv.ranks.size() + 1;
    }
    hdr.offPhfLevels = off; off = align256(off + lvDev.size() * sizeof(PhfLevelDev));
    hdr.offPhfBits = off; off = align256(off + phfWords * 8);
    hdr.offPhfRanks = off; off = align256(off + phfRanks * 8);
    hdr.offPhfFinal = off; off = align256(off + (h.phf.finalHash.size() + 1) * 16);
    hdr.offPhfData = off; off = align256(off + (h.phf.data.size() + 1) * 4);
    hdr.offPhfLens = off; off = align256(off + h.phf.lens.size() + 1);
    hdr.offPhfOverflow = off; off = align256(off + (h.phf.overflow.size() + 1) * 8);
  }
  uint64_t namesBytes = 0;
  for (const auto& nm : h.txpNames) namesBytes += nm.size() + 1;
  hdr.offNames = off; hdr.namesBytes = namesBytes; off = align256(off + namesBytes + 1);
  hdr.offSaTidPos = off; off = align256(off + n * 8);
  hdr.totalBytes = off;

  auto* idx = new rapmap_cuda_index();
  idx->device = device;
  idx->hdr = hdr;
  cudaError_t ce = cudaMalloc(&idx->blob, hdr.totalBytes);
  if (ce != cudaSuccess) { delete idx; return fail(RAPMAP_ERR_CUDA, std::string("cudaMalloc(index image): ") + cudaGetErrorString(ce)); }
  auto bail = [&](const std::string& m) { cudaFree(idx->blob); delete idx; return fail(RAPMAP_ERR_CUDA, m); };
#define IDX_TRY(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) return bail(std::string(#call) + ": " + cudaGetErrorString(e2)); } while (0)
  IDX_TRY(cudaMemcpy(idx->blob, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
  IDX_TRY(cudaMemcpy(idx->blob + hdr.offSA, h.SA.data(), n * 4, cudaMemcpyHostToDevice));
  IDX_TRY(cudaMemset(idx->blob + hdr.offText, 0, n + 256));
  IDX_TRY(cudaMemcpy(idx->blob + hdr.offText, h.text.data(), n, cudaMemcpyHostToDevice));
  {
    std::vector<uint4> rank(rankWords);
    uint32_t cum = 0;
    for (uint64_t w = 0; w < rankWords; ++w) {
      uint64_t bits = w < h.rsdBits.size() ? h.rsdBits[w] : 0;
      rank[w] = make_uint4(static_cast<uint32_t>(bits), static_cast<uint32_t>(bits >> 32), cum, 0);
      cum += static_cast<uint32_t>(__builtin_popcountll(bits));
    }
    IDX_TRY(cudaMemcpy(idx->blob + hdr.offRank, rank.data(), rankWords * 16, cudaMemcpyHostToDevice));
  }
  IDX_TRY(cudaMemcpy(idx->blob + hdr.offTxpOffsets, h.txpOffsets.data(), T * 4, cudaMemcpyHostToDevice));
  IDX_TRY(cudaMemcpy(idx->blob + hdr.offTxpLens, h.txpLens.data(), T * 4, cudaMemcpyHostToDevice));
  {
    std::string names;
    names.reserve(namesBytes);
    for (const auto& nm : h.txpNames) { names += nm; names.push_back('\0'); }
    IDX_TRY(cudaMemcpy(idx->blob + hdr.offNames, names.data(), names.size(), cudaMemcpyHostToDevice));
  }
  IDX_TRY(cudaMemset(idx->blob + hdr.offTable, 0xFF, slots * 16));
  if (h.perfectHash) {
    IDX_TRY(cudaMemset(idx->blob + hdr.offPhfLevels, 0, hdr.offNames - hdr.offPhfLevels));   // the PHF sections only: the names behind them are in place
    IDX_TRY(cudaMemcpy(idx->blob + hdr.offPhfLevels, lvDev.data(), lvDev.size() * sizeof(PhfLevelDev), cudaMemcpyHostToDevice));
    for (size_t i = 0; i < lvDev.size(); ++i) {
      const auto& lv = h.phf.levels[i];
      if (!lv.bits.empty()) IDX_TRY(cudaMemcpy(idx->blob + hdr.offPhfBits + lvDev[i].bitsOff * 8, lv.bits.data(), lv.bits.size() * 8, cudaMemcpyHostToDevice));
      if (!lv.ranks.empty()) IDX_TRY(cudaMemcpy(idx->blob + hdr.offPhfRanks + lvDev[i].ranksOff * 8, lv.ranks.data(), lv.ranks.size() * 8, cudaMemcpyHostToDevice));
    }
    if (!h.phf.finalHash.empty()) IDX_TRY(cudaMemcpy(idx->blob + hdr.offPhfFinal, h.phf.finalHash.data(), h.phf.finalHash.size() * 16, cudaMemcpyHostToDevice));
    if (!h.phf.data.empty()) IDX_TRY(cudaMemcpy(idx->blob + hdr.offPhfData, h.phf.data.data(), h.phf.data.size() * 4, cudaMemcpyHostToDevice));
    if (!h.phf.lens.empty()) IDX_TRY(cudaMemcpy(idx->blob + hdr.offPhfLens, h.phf.lens.data(), h.phf.lens.size(), cudaMemcpyHostToDevice));
    if (!h.phf.overflow.empty()) IDX_TRY(cudaMemcpy(idx->blob + hdr.offPhfOverflow, h.phf.overflow.data(), h.phf.overflow.size() * 8, cudaMemcpyHostToDevice));
  }
  if (deriveTable) {
    hdr.phfTableDerived = 0;   // the kernel below walks the BooPHF arrays
    derive_table_from_phf_kernel<<<4096, 256>>>(viewOf(idx->blob, hdr), reinterpret_cast<uint4*>(idx->blob + hdr.offTable), slots - 1);
    IDX_TRY(cudaDeviceSynchronize());
    hdr.phfTableDerived = 1;
    idx->hdr = hdr;
    IDX_TRY(cudaMemcpy(idx->blob, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
  }
  if (h.perfectHash && hdr.offFilter) {
    // -p index: the k-mer records are not stored, so the filter is filled from the text.  Every key FrugalBooMap::find can
    // return is verified against 31 text bases (include/FrugalBooMap.hpp:149-167), i.e. it is a k-mer of the text.
    uint32_t shift = 64;
    for (uint64_t w = hdr.filterWords; w > 1; w >>= 1) --shift;
    IDX_TRY(cudaMemset(idx->blob + hdr.offFilter, 0, hdr.filterWords * 4));
    build_filter_from_text_kernel<<<4096, 256>>>(idx->blob + hdr.offText, n, hdr.k, reinterpret_cast<uint32_t*>(idx->blob + hdr.offFilter), shift);
    IDX_TRY(cudaDeviceSynchronize());
  }
  // per SA entry: transcript and position inside it (the rank records and txpOffsets are in place above)
  build_sa_tidpos_kernel<<<4096, 256>>>(viewOf(idx->blob, hdr), reinterpret_cast<uint2*>(idx->blob + hdr.offSaTidPos));
  IDX_TRY(cudaDeviceSynchronize());
  {  // packed text
    uint32_t* dBad = nullptr;
    IDX_TRY(cudaMalloc(&dBad, 4));
    IDX_TRY(cudaMemset(dBad, 0, 4));
    build_text2_kernel<<<2048, 256>>>(idx->blob + hdr.offText, n, reinterpret_cast<TextRec*>(idx->blob + hdr.offText2), text2Recs, dBad);
    uint32_t bad = 0;
    cudaError_t e4 = cudaMemcpy(&bad, dBad, 4, cudaMemcpyDeviceToHost);
    cudaFree(dBad);
    if (e4 != cudaSuccess) return bail(std::string("packed text build: ") + cudaGetErrorString(e4));
    if (bad) {
      hdr.offText2 = 0;
      idx->hdr = hdr;
      IDX_TRY(cudaMemcpy(idx->blob, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
    }
  }
  if (!h.kmers.empty()) {
    KmerRecord* dRecs = nullptr;
    IDX_TRY(cudaMalloc(&dRecs, h.kmers.size() * sizeof(KmerRecord)));
    cudaError_t e3 = cudaMemcpy(dRecs, h.kmers.data(), h.kmers.size() * sizeof(KmerRecord), cudaMemcpyHostToDevice);
    if (e3 == cudaSuccess) {
      build_table_kernel<<<2048, 256>>>(dRecs, h.kmers.size(), reinterpret_cast<uint4*>(idx->blob + hdr.offTable), slots - 1);
      if (hdr.offFilter) {
        uint32_t shift = 64;
        for (uint64_t w = hdr.filterWords; w > 1; w >>= 1) --shift;
        e3 = cudaMemset(idx->blob + hdr.offFilter, 0, hdr.filterWords * 4);
        if (e3 == cudaSuccess) build_filter_kernel<<<2048, 256>>>(dRecs, h.kmers.size(), hdr.k, reinterpret_cast<uint32_t*>(idx->blob + hdr.offFilter), shift);
      }
      e3 = cudaDeviceSynchronize();
    }
    cudaFree(dRecs);
    if (e3 != cudaSuccess) return bail(std::string("hash table build: ") + cudaGetErrorString(e3));
  }
#undef IDX_TRY
  idx->view = viewOf(idx->blob, hdr);
  idx->names = std::move(h.txpNames);
  idx->lens = std::move(h.txpLens);
  *out = idx;
  return RAPMAP_OK;
}

int rapmap_cuda_index_load(const char* index_dir, int device, rapmap_cuda_index_t** out) {
  return guarded([&] { return indexLoadImpl(index_dir, device, out); });
}

void rapmap_cuda_index_free(rapmap_cuda_index_t* idx) {
  if (!idx) return;
  if (idx->ownsBlob && idx->blob) { cudaSetDevice(idx->device); cudaFree(idx->blob); }
  delete idx;
}

uint64_t rapmap_cuda_index_num_transcripts(const rapmap_cuda_index_t* idx) { return idx ? idx->hdr.numTxp : 0; }
const char* rapmap_cuda_index_transcript_name(const rapmap_cuda_index_t* idx, uint64_t tid) {
  return (idx && tid < idx->names.size()) ? idx->names[tid].c_str() : "";
}
uint64_t rapmap_cuda_index_transcript_len(const rapmap_cuda_index_t* idx, uint64_t tid) {
  return (idx && tid < idx->lens.size()) ? static_cast<uint64_t>(idx->lens[tid]) : 0;
}
uint32_t rapmap_cuda_index_k(const rapmap_cuda_index_t* idx) { return idx ? idx->hdr.k : 0; }
uint64_t rapmap_cuda_index_device_bytes(const rapmap_cuda_index_t* idx) { return idx ? idx->hdr.totalBytes : 0; }

int rapmap_cuda_index_image_bytes(const rapmap_cuda_index_t* idx, uint64_t* bytes) {
  if (!idx || !bytes) return fail(RAPMAP_ERR_ARG, "null argument");
  *bytes = idx->hdr.totalBytes;
  return RAPMAP_OK;
}
int rapmap_cuda_index_image_ptr(const rapmap_cuda_index_t* idx, void** p) {
  if (!idx || !p) return fail(RAPMAP_ERR_ARG, "null argument");
  *p = idx->blob;
  return RAPMAP_OK;
}
static int indexFromImageImpl(int device, void* blob, uint64_t bytes, rapmap_cuda_index_t** out) {
  if (!blob || !out) return fail(RAPMAP_ERR_ARG, "null argument");
  *out = nullptr;
  if (reinterpret_cast<uintptr_t>(blob) % 256 != 0) return fail(RAPMAP_ERR_ARG, "index image must be 256-byte aligned (the kernels use 256-bit loads)");
  if (bytes < sizeof(ImageHeader)) return fail(RAPMAP_ERR_ARG, "not an index image (too small)");
  CU_TRY(cudaSetDevice(device));
  ImageHeader hdr;
  CU_TRY(cudaMemcpy(&hdr, blob, sizeof(hdr), cudaMemcpyDeviceToHost));
  if (hdr.magic != kImageMagic || hdr.totalBytes != bytes) return fail(RAPMAP_ERR_ARG, "not an index image (bad magic or size)");
  {  // every section must lie inside the blob
    const uint64_t n = hdr.n, T = hdr.numTxp;
    struct { uint64_t off, len; } sec[] = {
      {hdr.offSA, n * 4}, {hdr.offText, n + 256}, {hdr.offRank, (n / 64 + 1) * 16}, {hdr.offTxpOffsets, T * 4}, {hdr.offTxpLens, T * 4},
      {hdr.offTable, hdr.tableSlots * 16}, {hdr.offFilter, hdr.offFilter ? hdr.filterWords * 4 : 0}, {hdr.offText2, hdr.offText2 ? (n / 32 + 2) * sizeof(TextRec) : 0},
      {hdr.offNames, hdr.namesBytes}, {hdr.offSaTidPos, n * 8}};
    for (const auto& sc : sec)
      if (sc.off > bytes || sc.len > bytes - sc.off) return fail(RAPMAP_ERR_ARG, "index image: a section lies outside the blob");
    if (hdr.tableSlots == 0 || (hdr.tableSlots & (hdr.tableSlots - 1)) != 0) return fail(RAPMAP_ERR_ARG, "index image: table size is not a power of two");
    if (hdr.hashKind && (hdr.offPhfLevels > bytes || hdr.offPhfOverflow > bytes)) return fail(RAPMAP_ERR_ARG, "index image: a section lies outside the blob");
  }
  auto* idx = new rapmap_cuda_index();
  idx->device = device;
  idx->blob = static_cast<uint8_t*>(blob);
  idx->ownsBlob = false;
  idx->hdr = hdr;
  idx->view = viewOf(idx->blob, hdr);
  // transcript names and lengths travel inside the image
  std::string names(hdr.namesBytes, '\0');
  idx->lens.resize(hdr.numTxp);
  cudaError_t e1 = hdr.namesBytes ? cudaMemcpy(&names[0], idx->blob + hdr.offNames, hdr.namesBytes, cudaMemcpyDeviceToHost) : cudaSuccess;
  cudaError_t e2 = hdr.numTxp ? cudaMemcpy(idx->lens.data(), idx->blob + hdr.offTxpLens, hdr.numTxp * 4, cudaMemcpyDeviceToHost) : cudaSuccess;
  if (e1 != cudaSuccess || e2 != cudaSuccess) { delete idx; return fail(RAPMAP_ERR_CUDA, "index image: reading the transcript table failed"); }
  idx->names.reserve(hdr.numTxp);
  for (size_t p = 0; p < names.size() && idx->names.size() < hdr.numTxp;) {
    const size_t e = names.find('\0', p);
    if (e == std::string::npos) break;
    idx->names.emplace_back(names, p, e - p);
    p = e + 1;
  }
  if (idx->names.size() != hdr.numTxp) { delete idx; return fail(RAPMAP_ERR_ARG, "index image: transcript name table is inconsistent"); }
  *out = idx;
  return RAPMAP_OK;
}

int rapmap_cuda_index_from_image(const rapmap_cuda_index_t* /*meta_src: unused since the image carries the names*/, int device, void* blob, uint64_t bytes,
                                 rapmap_cuda_index_t** out) {
  return guarded([&] { return indexFromImageImpl(device, blob, bytes, out); });
}

// -------------------------------------------------------------------------------------------------
static int deriveOpts(const rapmap_cuda_opts_t& o, DevOpts& d) {
  if (o.sel_aln) {
    // validateOpts, reference src/RapMapSAMapper.cpp:911-954
    if (o.consensus_slack < 0 || o.consensus_slack > 1) return fail(RAPMAP_ERR_ARG, "--consensusSlack must be between 0.0 and 1.0");
    if (o.min_score_fraction < 0 || o.min_score_fraction > 1) return fail(RAPMAP_ERR_ARG, "--minScoreFrac must be between 0.0 and 1.0");
    if (o.match_score <= 0) return fail(RAPMAP_ERR_ARG, "match score must be positive");
    if (o.mismatch_penalty > 0) return fail(RAPMAP_ERR_ARG, "mismatch penalty cannot be positive");
    int sv = 2 * (int(o.gap_open_penalty) + int(o.gap_extend_penalty)) + int(o.match_score);
    if (sv >= 127) return fail(RAPMAP_ERR_ARG, "[2*(gapOpen+gapExtend)+matchScore] cannot exceed 127");
  }
  if (o.max_mmp_extension < 1) return fail(RAPMAP_ERR_ARG, "--maxMMPExtension must be at least 1");
  std::memset(&d, 0, sizeof(d));
  d.maxNumHits = o.max_num_hits;
  d.covReq = o.quasi_coverage > 0.0 ? o.quasi_coverage : 0.0;
  d.maxInterval = 1000;
  d.doChaining = o.sel_aln;
  d.disableNIP = o.sensitive ? 1 : 0;      // reference src/RapMapSAMapper.cpp:386-388
  d.strictCheck = o.strict_check ? 1 : 0;  // :389
  d.considerMultiPos = o.sel_aln;
  d.selAln = o.sel_aln;
  d.fuzzy = o.fuzzy;
  d.consensusFraction = 1.0f;
  d.strictCheckSlack = 0;
  d.maxMMPExtension = 7;
  if (o.sel_aln) {  // reference src/RapMapSAMapper.cpp:410-417, include/SACollector.hpp:64,71
    d.consensusFraction = static_cast<float>((o.consensus_slack == 0.0) ? 1.0 : (1.0 - o.consensus_slack));
    d.strictCheckSlack = 1;
    d.maxMMPExtension = o.max_mmp_extension;
  }
  d.noOrphans = o.no_orphans;
  d.recoverOrphans = o.recover_orphans;  // acts only with the fuzzy merge (-s / -f), src/RapMapSAMapper.cpp:498
  d.noDovetail = o.no_dovetail;
  d.hardFilter = o.hard_filter;
  d.alignmentPolicy = o.alignment_policy;
  d.ma = o.match_score; d.mm = o.mismatch_penalty; d.go = o.gap_open_penalty; d.ge = o.gap_extend_penalty;
  d.dpBandwidth = o.dp_bandwidth;
  d.minScoreFraction = o.min_score_fraction;
  return RAPMAP_OK;
}

static void freeMapperBuffers(rapmap_cuda_mapper* m) {
  for (auto& sl : m->slots) {
    for (int i = 0; i < 2; ++i) { cudaFree(sl.dSeq[i]); cudaFree(sl.dOff[i]); }
    cudaFree(sl.dPairOff); cudaFree(sl.dHits); cudaFree(sl.dSelOut);
    if (sl.hStage) cudaFreeHost(sl.hStage);
    for (auto& e : sl.ev) if (e) cudaEventDestroy(e);
    if (sl.evIn) cudaEventDestroy(sl.evIn);
    if (sl.evCompute) cudaEventDestroy(sl.evCompute);
    if (sl.evOut) cudaEventDestroy(sl.evOut);
  }
  cudaFree(m->dSumm); cudaFree(m->dIvArena); cudaFree(m->dQSumm); cudaFree(m->dQaArena); cudaFree(m->dPosPool);
  cudaFree(m->dScratch); cudaFree(m->dPairCount); cudaFree(m->dCubTemp);
  cudaFree(m->dCtl); cudaFree(m->dCounters);
  cudaFree(m->dPacked); cudaFree(m->dKmask); cudaFree(m->dOrder); cudaFree(m->dClassCtl); cudaFree(m->dOrder2); cudaFree(m->dK2Class); cudaFree(m->dIvScratch); cudaFree(m->dVoteScratch);
  selAlnFree(m->selaln);
  if (m->stream) cudaStreamDestroy(m->stream);
  if (m->sIn) cudaStreamDestroy(m->sIn);
  if (m->sOut) cudaStreamDestroy(m->sOut);
}

// RAPMAP_B200_TINY_ARENAS=1 (tests only): every growable device work area starts far too small, so that the first batch
// of a mapper walks through each overflow -> grow -> re-run path of finishBatch().
// Copy-engine part of the copy-out (enqueueAttempt): by default for batches of 32k pairs and more (below that two more DMA
// submissions cost more than the kernel's PCIe stores).  RAPMAP_B200_COPYOUT=kernel: never (A/B); =engine: for every batch
// (tests: small batches through the same path).  Read per batch, so a test can switch it.
static bool speculativeCopy(uint64_t n) {
  const char* t = std::getenv("RAPMAP_B200_COPYOUT");
  if (t && std::string(t) == "kernel") return false;
  if (t && std::string(t) == "engine") return true;
  return n >= 32768;
}

static bool tinyArenas() {
  const char* t = std::getenv("RAPMAP_B200_TINY_ARENAS");
  return t && t[0] == '1';
}

static int mapperCreateImpl(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, uint64_t max_batch, uint32_t max_read_len,
                            rapmap_cuda_mapper_t** out) {
  if (!idx || !opts || !out || max_batch == 0) return fail(RAPMAP_ERR_ARG, "null / zero argument");
  *out = nullptr;
  if (max_read_len < idx->hdr.k || max_read_len > 1000) return fail(RAPMAP_ERR_ARG, "max_read_len must be in [k, 1000]");
  if (max_batch > (1ull << 30)) return fail(RAPMAP_ERR_ARG, "max_batch too large");
  DevOpts d;
  int rc = deriveOpts(*opts, d);
  if (rc) return rc;
  CU_TRY(cudaSetDevice(idx->device));
  auto* m = new rapmap_cuda_mapper();
  m->idx = idx; m->opts = *opts; m->dopts = d; m->maxBatch = max_batch; m->maxReadLen = max_read_len;
  auto bail = [&](const std::string& msg) { freeMapperBuffers(m); delete m; return fail(RAPMAP_ERR_CUDA, msg); };
#define M_TRY(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) return bail(std::string(#call) + ": " + cudaGetErrorString(e2)); } while (0)
  const bool tiny = tinyArenas();
  const bool needPos = opts->sel_aln || opts->fuzzy;
  cudaDeviceProp prop;
  M_TRY(cudaGetDeviceProperties(&prop, idx->device));
  m->numSMs = prop.multiProcessorCount;
  M_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  M_TRY(cudaStreamCreateWithFlags(&m->sIn, cudaStreamNonBlocking));
  M_TRY(cudaStreamCreateWithFlags(&m->sOut, cudaStreamNonBlocking));
  const uint64_t R = 2 * max_batch;
  m->seqCap = max_batch * max_read_len;
  m->hitsCap = tiny ? 8 : max_batch * 6 + 1024;
  for (auto& sl : m->slots) {
    for (auto& e : sl.ev) M_TRY(cudaEventCreate(&e));
    M_TRY(cudaEventCreateWithFlags(&sl.evIn, cudaEventDisableTiming));
    M_TRY(cudaEventCreateWithFlags(&sl.evCompute, cudaEventDisableTiming));
    M_TRY(cudaEventCreateWithFlags(&sl.evOut, cudaEventBlockingSync | cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      M_TRY(cudaMalloc(&sl.dSeq[i], m->seqCap + 16));
      M_TRY(cudaMalloc(&sl.dOff[i], (max_batch + 1) * 8));
    }
    M_TRY(cudaMalloc(&sl.dPairOff, (max_batch + 1) * 8));
    M_TRY(cudaMalloc(&sl.dHits, m->hitsCap * sizeof(rapmap_hit_t)));
    if (opts->sel_aln) M_TRY(cudaMalloc(&sl.dSelOut, m->hitsCap * sizeof(rapmap_hit_t)));
    M_TRY(cudaMallocHost(&sl.hStage, sizeof(StageBlock)));
  }
  M_TRY(cudaMalloc(&m->dSumm, R * sizeof(ReadSummary)));
  M_TRY(cudaMalloc(&m->dQSumm, R * sizeof(QASummary)));
  m->qaCap = tiny ? 8u : static_cast<uint32_t>(std::min<uint64_t>(R * 8 + 1024, 0xFFFFFFF0ull));
  M_TRY(cudaMalloc(&m->dQaArena, static_cast<uint64_t>(m->qaCap) * sizeof(QARec)));
  // position lists exist whenever the fuzzy merge runs (-s or -f)
  m->posCap = (needPos && !tiny) ? static_cast<uint32_t>(std::min<uint64_t>(R * 12 + 1024, 0xFFFFFFF0ull)) : 16;
  M_TRY(cudaMalloc(&m->dPosPool, static_cast<uint64_t>(m->posCap) * 4));
  M_TRY(cudaMalloc(&m->dPairCount, (max_batch + 1) * 4));
  M_TRY(cudaMalloc(&m->dCtl, 8 * 4));
  M_TRY(cudaMalloc(&m->dCounters, sizeof(Counters5)));
  {
    cub::TransformInputIterator<uint64_t, CastU64, uint32_t*> it(m->dPairCount, CastU64());
    M_TRY(cub::DeviceScan::ExclusiveSum(nullptr, m->cubTempBytes, it, m->slots[0].dPairOff, static_cast<int>(max_batch + 1)));
    M_TRY(cudaMalloc(&m->dCubTemp, m->cubTempBytes + 16));
  }
  // ---- launch geometry: persistent grids, whole multiples of the SM count
  m->pmax = max_read_len - idx->hdr.k + 1;
  int occ = 0;
  // NB: no cudaSharedmemCarveoutMaxShared here - these kernels live off L1 hits on the text / SA / table sectors; forcing the
  // maximum shared-memory carve-out shrank L1 and cost 20 %.
  {  // SA-lookup kernel: one thread per read
    m->laneWords = (max_read_len + 31) / 32;
    m->laneSmem = 2u * m->laneWords * 16u * kLaneThreads;  // packed read words + k-mer mask words
    if (m->laneSmem > 227 * 1024) { m->laneSmem /= 2; m->masksInGlobal = true; }  // very long reads: masks stay in global memory
    if (m->laneSmem > 227 * 1024) return bail("max_read_len too large for the shared-memory read words");
    const bool general = !(d.disableNIP && d.strictCheck);
    m->laneKernel = idx->view.hashKind ? (general ? &sa_collect_lane_kernel<kLaneThreads, kLaneMinBlocks, true, true> : &sa_collect_lane_kernel<kLaneThreads, kLaneMinBlocks, true, false>)
                                      : (general ? &sa_collect_lane_kernel<kLaneThreads, kLaneMinBlocks, false, true> : &sa_collect_lane_kernel<kLaneThreads, kLaneMinBlocks, false, false>);
    M_TRY(cudaFuncSetAttribute(m->laneKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(m->laneSmem)));
    M_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, m->laneKernel, kLaneThreads, m->laneSmem));
    if (occ < 1) return bail("sa_collect_lane_kernel does not fit on an SM");
    m->gridLane = m->numSMs * occ;
    m->scratchSlotsK1 = static_cast<uint64_t>(m->gridLane) * kLaneThreads;
    M_TRY(cudaMalloc(&m->dPacked, R * m->laneWords * sizeof(uint4)));
    M_TRY(cudaMalloc(&m->dKmask, R * m->laneWords * sizeof(uint4)));
    M_TRY(cudaMemset(m->dKmask, 0, R * m->laneWords * sizeof(uint4)));
    M_TRY(cudaMalloc(&m->dOrder, R * 4));
    M_TRY(cudaMalloc(&m->dClassCtl, 2 * kWorkClasses * 4));
    // interval arena: 4 (10 with chaining) records per read, plus the unused tails of the RAPMAP_LANE_CHUNK-record slices
    // every resident warp reserves
    const uint64_t want = R * (needPos ? 10 : 4) + 1024 + static_cast<uint64_t>(m->gridLane) * (kLaneThreads / 32) * RAPMAP_LANE_CHUNK;
    m->ivCap = tiny ? 64u : static_cast<uint32_t>(std::min<uint64_t>(want, 0xFFFFFFF0ull));
    M_TRY(cudaMalloc(&m->dIvArena, static_cast<uint64_t>(m->ivCap) * sizeof(IntervalRec)));
    m->ivStride = tiny ? 1u : std::min<uint32_t>(m->pmax, 24);
    M_TRY(cudaMalloc(&m->dIvScratch, m->scratchSlotsK1 * 2 * m->ivStride * sizeof(IntervalRec)));
    const bool voteMode = d.strictCheck && !(d.disableNIP && d.strictCheck);
    if (voteMode) {
      m->voteWords = (m->pmax + 31) / 32;
      M_TRY(cudaMalloc(&m->dVoteScratch, m->scratchSlotsK1 * 3 * m->voteWords * 4));
    }
  }
  {  // lane-per-read hit resolution: plain form for default quasimap, chaining form for -s / -f; RAPMAP_B200_K2=warp turns both off
    const char* sel = std::getenv("RAPMAP_B200_K2");
    const bool off = sel && std::string(sel) == "warp";
    m->laneMap = !d.selAln && !d.fuzzy && !d.doChaining && !off;
    m->chainLaneMap = !m->laneMap && !off;
    if (m->chainLaneMap) {
      int ci = 0;
#define SETUP_CHAIN(NT_, CAP_)                                                                                                                       \
      {                                                                                                                                              \
        const uint32_t smemC = chainLaneStride(CAP_) * NT_;                                                                                          \
        M_TRY(cudaFuncSetAttribute(hits_to_mappings_chain_lane_kernel<NT_, CAP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemC))); \
        M_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hits_to_mappings_chain_lane_kernel<NT_, CAP_>, NT_, smemC));                         \
        if (occ < 1) return bail("hits_to_mappings_chain_lane_kernel does not fit on an SM");                                                         \
        m->gridChain[ci++] = m->numSMs * occ;                                                                                                          \
      }
      RAPMAP_CHAIN_CLASSES(SETUP_CHAIN)
#undef SETUP_CHAIN
      M_TRY(cudaMalloc(&m->dOrder2, R * 4));
      M_TRY(cudaMalloc(&m->dK2Class, 2 * kK2Buckets * 4));
    }
    if (m->laneMap) {
      M_TRY(cudaFuncSetAttribute(hits_to_mappings_lane_kernel<kMapLaneThreads, kMapLaneCap>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMapLaneSmem)));
      M_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hits_to_mappings_lane_kernel<kMapLaneThreads, kMapLaneCap>, kMapLaneThreads, kMapLaneSmem));
      if (occ < 1) return bail("hits_to_mappings_lane_kernel does not fit on an SM");
      m->gridLaneMap = m->numSMs * occ;
    }
  }
  m->mapSmem = static_cast<uint32_t>(workAreaBytes(m->smemEntries)) * kWarps;
  M_TRY(cudaFuncSetAttribute(hits_to_mappings_kernel<kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(m->mapSmem)));
  M_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hits_to_mappings_kernel<kWarps>, kWarps * 32, m->mapSmem));
  if (occ < 1) return bail("hits_to_mappings_kernel does not fit on an SM");
  m->gridMap = m->numSMs * occ;
  // global work strip per resident warp: the worst case of one strand is (pmax intervals) x (maxInterval-1 entries);
  // 8192 entries cover every read whose expanded intervals total <= 4096 SA entries (pow2 padding), larger reads are
  // reported through kStatScratchFull and re-run with a strip sized from the actual maximum.
  m->scratchEntries = tiny ? 128 : 8192;
  m->scratchStride = workAreaBytes(m->scratchEntries);
  M_TRY(cudaMalloc(&m->dScratch, m->scratchStride * static_cast<uint64_t>(m->gridMap) * kWarps));
  if (opts->sel_aln) {
    cudaError_t e2 = selAlnAlloc(m->selaln, max_batch, max_read_len, m->hitsCap);
    if (e2 != cudaSuccess) return bail(std::string("selAlnAlloc: ") + cudaGetErrorString(e2));
    std::string serr;
    int rc2 = selAlnSetup(m->selaln, m->selLaunch, serr);
    if (rc2) { freeMapperBuffers(m); delete m; return fail(rc2, serr); }
  }
#undef M_TRY
  *out = m;
  return RAPMAP_OK;
}

static int growU32(void** p, uint32_t& cap, uint64_t need, size_t elem) {
  uint64_t nc = need + need / 4 + 1024;
  if (nc > 0xFFFFFFF0ull) return fail(RAPMAP_ERR_CAPACITY, "device work arena would exceed 2^32 records; use smaller batches");
  cudaFree(*p);
  *p = nullptr;
  CU_TRY(cudaMalloc(p, nc * elem));
  cap = static_cast<uint32_t>(nc);
  return RAPMAP_OK;
}

// Enqueues one attempt at a batch: K0 pack / k-mer masks / work classes, K1 SA lookup, K2 hit resolution, K3 merge
// (count -> scan -> write), K4 selective alignment and the copy of the control block to pinned host memory on the compute
// stream, then the result copy-out on the output stream.  Nothing here waits for the device: every kernel launches
// against the CURRENT capacities and raises a status bit instead of writing past them (finishBatch grows what overflowed
// and calls this again).
static int enqueueAttempt(rapmap_cuda_mapper* m, BatchSlot& sl) {
  cudaStream_t st = m->stream;
  const BatchView& bv = sl.view;
  const uint64_t n = bv.n;
  const bool paired = sl.paired;
  CU_TRY(cudaStreamWaitEvent(st, sl.evIn, 0));
  CU_TRY(cudaMemsetAsync(m->dCtl, 0, 32, st));
  CU_TRY(cudaMemsetAsync(m->dCounters, 0, sizeof(Counters5), st));
  CU_TRY(cudaEventRecord(sl.ev[11], st));
  // ---- kernel 1: SA lookup
  LaneParams lp{};
  lp.ix = m->idx->view; lp.reads = bv; lp.opts = m->dopts; lp.maxReadLen = m->maxReadLen; lp.nw = m->laneWords; lp.packed = m->dPacked;
  lp.kmask = m->dKmask; lp.maskChunks = (m->pmax + 7) / 8; lp.classCtl = m->dClassCtl; lp.order = m->dOrder; lp.masksInGlobal = m->masksInGlobal ? 1u : 0u;
  lp.summ = m->dSumm; lp.arena = m->dIvArena; lp.arenaCap = m->ivCap; lp.arenaCursor = m->dCtl + 0; lp.status = m->dCtl + 3;
  lp.ivScratch = m->dIvScratch; lp.ivStride = m->ivStride; lp.voteScratch = m->dVoteScratch; lp.voteWords = m->voteWords; lp.readCursor = m->dCtl + 4;
  const int g0 = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(m->numSMs) * 8, (bv.numReads * m->laneWords + 255) / 256));
  pack_reads_kernel<<<g0, 256, 0, st>>>(lp);
  const int g0b = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(m->numSMs) * 8, (bv.numReads * lp.maskChunks + 255) / 256));
  kmer_mask_kernel<<<g0b, 256, 0, st>>>(lp);
  CU_TRY(cudaMemsetAsync(m->dClassCtl, 0, 2 * kWorkClasses * 4, st));
  const int g0c = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(m->numSMs) * 8, (bv.numReads + 255) / 256));
  work_class_hist_kernel<<<g0c, 256, 0, st>>>(lp);
  work_class_scatter_kernel<<<g0c, 256, 0, st>>>(lp);
  CU_TRY(cudaEventRecord(sl.ev[8], st));
  const int g1 = static_cast<int>(std::min<uint64_t>(m->gridLane, (bv.numReads + kLaneThreads - 1) / kLaneThreads));
  m->laneKernel<<<g1, kLaneThreads, m->laneSmem, st>>>(lp);
  sl.launches += 5;
  CU_TRY(cudaEventRecord(sl.ev[2], st));
  // ---- kernel 2: hit resolution
  MapParams mp{};
  mp.ix = m->idx->view; mp.opts = m->dopts; mp.numReads = bv.numReads; mp.numPairs = n; mp.pairedInput = paired ? 1 : 0;
  mp.summ = m->dSumm; mp.arena = m->dIvArena; mp.qsumm = m->dQSumm; mp.qaArena = m->dQaArena; mp.qaCap = m->qaCap; mp.qaCursor = m->dCtl + 1;
  mp.posPool = m->dPosPool; mp.posCap = m->posCap; mp.posCursor = m->dCtl + 2;
  mp.scratch = m->dScratch; mp.scratchEntries = m->scratchEntries; mp.scratchStride = m->scratchStride; mp.smemEntries = m->smemEntries;
  mp.status = m->dCtl + 3;
  if (m->chainLaneMap) {
    // -s / -f: reads grouped by their number of SA entries, then thread-per-read with chaining and position lists, one launch
    // per size range (strip and block size to match); what is bigger than the biggest strip goes to the warp-per-read kernel
    CU_TRY(cudaMemsetAsync(m->dK2Class, 0, 2 * kK2Buckets * 4, st));
    const int gc = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(m->numSMs) * 8, (bv.numReads + 255) / 256));
    mp.classHist = m->dK2Class;
    k2_class_hist_kernel<<<gc, 256, 0, st>>>(mp, kChainLaneMaxIv);
    k2_class_scatter_kernel<<<gc, 256, 0, st>>>(mp, kChainLaneMaxIv, m->dOrder2);
    sl.launches += 2;
    mp.order = m->dOrder2;
    int ci = 0;
    uint32_t lo = 1;
#define LAUNCH_CHAIN(NT_, CAP_)                                                                                                              \
    {                                                                                                                                        \
      mp.bLo = lo; mp.bHi = CAP_; lo = CAP_ + 1;                                                                                             \
      const int gl = static_cast<int>(std::min<uint64_t>(m->gridChain[ci++], (bv.numReads + NT_ - 1) / NT_));                                  \
      hits_to_mappings_chain_lane_kernel<NT_, CAP_><<<gl, NT_, chainLaneStride(CAP_) * NT_, st>>>(mp);                                         \
      ++sl.launches;                                                                                                                         \
    }
    RAPMAP_CHAIN_CLASSES(LAUNCH_CHAIN)
#undef LAUNCH_CHAIN
    mp.bLo = kK2MaxEntries + 1; mp.bHi = kK2MaxEntries + 1;  // the warp-per-read kernel below takes the last bucket
  }
  if (m->laneMap) {  // small reads thread-per-read; the rest (marked) by the warp-per-read kernel below
    const int gl = static_cast<int>(std::min<uint64_t>(m->gridLaneMap, (bv.numReads + kMapLaneThreads - 1) / kMapLaneThreads));
    hits_to_mappings_lane_kernel<kMapLaneThreads, kMapLaneCap><<<gl, kMapLaneThreads, kMapLaneSmem, st>>>(mp);
    ++sl.launches;
    mp.skipDone = 1;
  }
  const int g2 = static_cast<int>(std::min<uint64_t>(m->gridMap, (bv.numReads + kWarps - 1) / kWarps));
  hits_to_mappings_kernel<kWarps><<<g2, kWarps * 32, m->mapSmem, st>>>(mp);
  ++sl.launches;
  CU_TRY(cudaEventRecord(sl.ev[3], st));
  // ---- kernel 3: mate merge: count -> exclusive scan -> write at the final, input-ordered offsets
  MergeParams gp{};
  gp.opts = m->dopts; gp.numPairs = n; gp.pairedInput = paired ? 1 : 0; gp.qsumm = m->dQSumm; gp.qaArena = m->dQaArena; gp.summ = m->dSumm;
  gp.pairCount = m->dPairCount; gp.pairOffset = sl.dPairOff; gp.hits = sl.dHits; gp.hitsCap = m->hitsCap; gp.counters = m->dCounters;
  gp.posPool = m->dPosPool;
  gp.reads = bv; gp.text = m->idx->view.text; gp.txpOffsets = m->idx->view.txpOffsets; gp.txpLens = m->idx->view.txpLens;
  const int g3 = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(m->numSMs) * 8, (n + 255) / 256));
  const bool fuzzy = paired && (m->dopts.selAln || m->dopts.fuzzy);  // unmated reads have no mate merge (processReadsSingleSA)
  CU_TRY(cudaMemsetAsync(m->dPairCount + n, 0, 4, st));
  if (fuzzy) merge_count_kernel<true><<<g3, 256, 0, st>>>(gp);
  else merge_count_kernel<false><<<g3, 256, 0, st>>>(gp);
  ++sl.launches;
  {
    cub::TransformInputIterator<uint64_t, CastU64, uint32_t*> it(m->dPairCount, CastU64());
    size_t tb = m->cubTempBytes;
    CU_TRY(cub::DeviceScan::ExclusiveSum(m->dCubTemp, tb, it, sl.dPairOff, static_cast<int>(n + 1), st));
  }
  CU_TRY(cudaMemcpyAsync(&sl.hStage->mergeTotal, sl.dPairOff + n, 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaEventRecord(sl.ev[4], st));
  // optimistic: written against hitsCap (a pair whose slice would pass it is skipped; the host sees mergeTotal > hitsCap)
  if (fuzzy) merge_write_kernel<true><<<g3, 256, 0, st>>>(gp);
  else merge_write_kernel<false><<<g3, 256, 0, st>>>(gp);
  ++sl.launches;
  CU_TRY(cudaEventRecord(sl.ev[5], st));
  // ---- selective alignment (ksw2 scoring + score filter): survivors to the slot's dSelOut, dPairOff rewritten in place
  sl.hStage->selTotal = 0; sl.hStage->dpJobs[0] = 0; sl.hStage->dpJobs[1] = 0; sl.hStage->dpJobs[2] = 0; sl.hStage->dpJobs[3] = 0;
  if (m->dopts.selAln) {
    int rc = selAlnEnqueue(m->selaln, m->selLaunch, m->idx->view, m->dopts, bv, n, paired, sl.dHits, sl.dSelOut, sl.dPairOff, m->hitsCap, m->dCubTemp,
                           m->cubTempBytes, m->numSMs, st, &sl.launches, &sl.hStage->selTotal, sl.hStage->dpJobs, sl.ev[9], sl.ev[10], g_err);
    if (rc) return rc;
  }
  CU_TRY(cudaEventRecord(sl.ev[6], st));
  CU_TRY(cudaMemcpyAsync(sl.hStage->ctl, m->dCtl, 16, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&sl.hStage->counters, m->dCounters, sizeof(Counters5), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaEventRecord(sl.evCompute, st));
  // ---- results out, on their own stream: the next batch computes meanwhile
  CU_TRY(cudaStreamWaitEvent(m->sOut, sl.evCompute, 0));
  if (sl.directOut) {
    const rapmap_hit_t* src = m->dopts.selAln ? sl.dSelOut : sl.dHits;
    // a small grid for PCIe: ~16k threads with a 16-byte store each cover the link's bandwidth-delay product and leave the
    // SMs to the next batch's kernels
    const int gco = sl.out->location == RAPMAP_LOC_DEVICE ? m->numSMs * 4 : 64;
    // Pinned host buffers: the copy engines move the offsets (fixed size) and the part of the records the previous batches
    // make certain enough (98 % of their records-per-pair rate), the kernel only the tail whose length is known on the
    // device alone.  A kernel that pushes the whole 100+ MB over PCIe keeps its SMs' memory pipes full of posted writes for
    // 2 ms and doubled the time of the next batch's first kernels (profiles/r02j_e2e_diag.txt); records past num_hits that
    // the speculative copy may bring along lie inside hits_capacity and are unspecified anyway.
    uint64_t skip = 0;
    const uint64_t* offSrc = sl.dPairOff;
    if (sl.out->location == RAPMAP_LOC_HOST && speculativeCopy(n)) {
      CU_TRY(cudaMemcpyAsync(sl.out->pair_offsets, sl.dPairOff, (n + 1) * 8, cudaMemcpyDeviceToHost, m->sOut));
      offSrc = nullptr;
      if (sl.outHitsDev != nullptr && m->hitsPerPair > 0.0) {
        skip = static_cast<uint64_t>(m->hitsPerPair * 0.98 * static_cast<double>(n));
        skip = std::min<uint64_t>(skip, std::min<uint64_t>(sl.out->hits_capacity, m->hitsCap)) & ~3ULL;
        if (skip > 0) CU_TRY(cudaMemcpyAsync(sl.out->hits, src, skip * sizeof(rapmap_hit_t), cudaMemcpyDeviceToHost, m->sOut));
      }
    }
    copy_out_kernel<<<gco, 256, 0, m->sOut>>>(src, static_cast<rapmap_hit_t*>(sl.outHitsDev), sl.out->hits_capacity, sl.dPairOff + n, offSrc,
                                               static_cast<uint64_t*>(sl.outOffDev), n + 1, skip);
    ++sl.launches;
  }
  CU_TRY(cudaEventRecord(sl.ev[7], m->sOut));
  CU_TRY(cudaEventRecord(sl.evOut, m->sOut));
  return RAPMAP_OK;
}

// Waiting for a batch.  Small batches spin (cudaStreamSynchronize: a wake-up latency would show).  Big batches poll
// cudaEventQuery with 50 us naps: the thread uses next to no CPU (several ranks per box: spinning threads fight the launching
// ones for cores) and there is no interrupt in the path - sleeping on a cudaEventBlockingSync event cost a quarter of
// the resident rate with 8 ranks on one virtualised box (1041 vs 1450 M pairs/s, profiles/r02k_n8_diag.txt).
// RAPMAP_B200_WAIT=block selects the blocking event.
static int waitMode() {
  static const int v = [] { const char* t = std::getenv("RAPMAP_B200_WAIT"); return (t && std::string(t) == "block") ? 0 : 1; }();
  return v;
}

static cudaError_t waitSlot(rapmap_cuda_mapper* m, BatchSlot& sl) {
  if (sl.view.n < 32768) return cudaStreamSynchronize(m->sOut);
  if (waitMode() == 1) {
    for (;;) {
      const cudaError_t e = cudaEventQuery(sl.evOut);
      if (e != cudaErrorNotReady) return e;
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  }
  return cudaEventSynchronize(sl.evOut);
}

static int mapBatchAsyncImpl(rapmap_cuda_mapper_t* m, const rapmap_read_batch_t* reads, rapmap_hit_batch_t* out) {
  if (!m || !reads || !out) return fail(RAPMAP_ERR_ARG, "null argument");
  if (m->submitted - m->collected >= static_cast<uint64_t>(kDepth))
    return fail(RAPMAP_ERR_ARG, "the mapper already has its maximum of batches in flight (call rapmap_cuda_mapper_wait first)");
  BatchSlot& sl = m->slots[m->submitted % kDepth];
  sl.out = out;
  sl.rerun = false;
  sl.launches = 0;
  if (reads->n == 0) {
    sl.view = BatchView{};
    sl.inFlight = true;
    ++m->submitted;
    return RAPMAP_OK;
  }
  if (reads->n > m->maxBatch) return fail(RAPMAP_ERR_ARG, "batch larger than the mapper's max_batch");
  if (!reads->seq1) return fail(RAPMAP_ERR_ARG, "seq1 is null");
  if (!reads->off1 && (reads->fixed_len == 0 || reads->fixed_len > m->maxReadLen)) return fail(RAPMAP_ERR_ARG, "fixed_len must be in [1, max_read_len]");
  if (reads->seq2 && ((reads->off1 == nullptr) != (reads->off2 == nullptr))) return fail(RAPMAP_ERR_ARG, "off1/off2 must both be set or both be null");
  if (!out->pair_offsets) return fail(RAPMAP_ERR_ARG, "pair_offsets is null");
  CU_TRY(cudaSetDevice(m->idx->device));
  const uint64_t n = reads->n;
  sl.paired = reads->seq2 != nullptr;

  // ---- reads to the device, on the input stream
  CU_TRY(cudaEventRecord(sl.ev[0], m->sIn));
  BatchView bv{};
  bv.n = n; bv.numReads = sl.paired ? 2 * n : n; bv.fixedLen = reads->fixed_len;
  for (int mate = 0; mate < (sl.paired ? 2 : 1); ++mate) {
    const uint8_t* seq = mate ? reads->seq2 : reads->seq1;
    const uint64_t* off = mate ? reads->off2 : reads->off1;
    if (reads->location == RAPMAP_LOC_DEVICE) {
      bv.seq[mate] = seq; bv.off[mate] = off;
    } else {
      uint64_t bytes = off ? off[n] - off[0] : n * reads->fixed_len;
      if (off && off[0] != 0) return fail(RAPMAP_ERR_ARG, "offsets must start at 0");
      if (bytes > m->seqCap) return fail(RAPMAP_ERR_ARG, "read bases exceed max_batch * max_read_len");
      CU_TRY(cudaMemcpyAsync(sl.dSeq[mate], seq, bytes, cudaMemcpyHostToDevice, m->sIn));
      bv.seq[mate] = sl.dSeq[mate];
      if (off) { CU_TRY(cudaMemcpyAsync(sl.dOff[mate], off, (n + 1) * 8, cudaMemcpyHostToDevice, m->sIn)); bv.off[mate] = sl.dOff[mate]; }
      else bv.off[mate] = nullptr;
    }
  }
  sl.view = bv;
  CU_TRY(cudaEventRecord(sl.ev[1], m->sIn));
  CU_TRY(cudaEventRecord(sl.evIn, m->sIn));
  sl.outHitsDev = deviceAlias(out->hits, out->location);
  sl.outOffDev = deviceAlias(out->pair_offsets, out->location);
  sl.directOut = sl.outOffDev != nullptr && (sl.outHitsDev != nullptr || out->hits == nullptr || out->hits_capacity == 0) &&
                 (reinterpret_cast<uintptr_t>(sl.outOffDev) & 7) == 0 && (reinterpret_cast<uintptr_t>(sl.outHitsDev) & 3) == 0;
  int rc = enqueueAttempt(m, sl);
  if (rc) return rc;
  sl.inFlight = true;
  ++m->submitted;
  return RAPMAP_OK;
}

// Waits for the OLDEST batch in flight; grows what overflowed and re-runs (deterministic: same batch, bigger arenas).  With
// device or pinned host output buffers the attempt has already moved the result out (copy_out_kernel): ONE host
// synchronisation per batch.  Pageable host buffers take a cudaMemcpyAsync of the now known size and a second one.
static int mapperWaitImpl(rapmap_cuda_mapper_t* m) {
  if (!m) return fail(RAPMAP_ERR_ARG, "null argument");
  if (m->submitted == m->collected) return fail(RAPMAP_ERR_ARG, "no batch in flight");
  BatchSlot& sl = m->slots[m->collected % kDepth];
  ++m->collected;
  sl.inFlight = false;
  rapmap_hit_batch_t* out = sl.out;
  std::memset(&m->timing, 0, sizeof(m->timing));
  const uint64_t n = sl.view.n;
  m->lastReads = sl.view.numReads;
  if (n == 0) {
    out->num_hits = 0;
    std::memset(out->counters, 0, sizeof(out->counters));
    if (out->pair_offsets && out->location == RAPMAP_LOC_HOST) out->pair_offsets[0] = 0;
    return RAPMAP_OK;
  }
  CU_TRY(cudaSetDevice(m->idx->device));
  uint32_t retries = 0;
  uint64_t total = 0;
  for (;; ++retries) {
    CU_TRY(waitSlot(m, sl));
    CU_TRY(cudaGetLastError());
    if (retries > 10) return fail(RAPMAP_ERR_CAPACITY, "device work arenas kept overflowing");
    if (sl.rerun) {  // an earlier batch's overflow re-allocated arenas while this one was queued behind it
      sl.rerun = false;
      int rc = enqueueAttempt(m, sl);
      if (rc) return rc;
      continue;
    }
    const uint32_t status = sl.hStage->ctl[3];
    if (status & kStatReadTooLong) return fail(RAPMAP_ERR_ARG, "a read is longer than the mapper's max_read_len");
    if (status & kStatInternal) return fail(RAPMAP_ERR_CUDA, "internal error: a kernel met a case its launch configuration excludes");
    const bool hitsFull = sl.hStage->mergeTotal > m->hitsCap;
    if ((status & (kStatIntervalArenaFull | kStatIvScratchFull | kStatQAArenaFull | kStatPosPoolFull | kStatScratchFull)) == 0 && !hitsFull) break;
    // ---- something overflowed: drain the device (a later batch may be running on the arenas), grow, repeat the attempt
    CU_TRY(cudaStreamSynchronize(m->stream));
    CU_TRY(cudaStreamSynchronize(m->sOut));
    for (auto& other : m->slots)
      if (&other != &sl && other.inFlight) other.rerun = true;
    if (status & kStatIntervalArenaFull) { int rc = growU32(reinterpret_cast<void**>(&m->dIvArena), m->ivCap, sl.hStage->ctl[0], sizeof(IntervalRec)); if (rc) return rc; }
    if (status & kStatIvScratchFull) {  // a read produced more intervals per strand than the per-thread list holds: size it for the worst case
      if (m->ivStride >= m->pmax) return fail(RAPMAP_ERR_CAPACITY, "interval scratch overflow at worst-case size");
      cudaFree(m->dIvScratch); m->dIvScratch = nullptr;
      m->ivStride = m->pmax;
      CU_TRY(cudaMalloc(&m->dIvScratch, m->scratchSlotsK1 * 2 * m->ivStride * sizeof(IntervalRec)));
    }
    if (status & kStatQAArenaFull) { int rc = growU32(reinterpret_cast<void**>(&m->dQaArena), m->qaCap, sl.hStage->ctl[1], sizeof(QARec)); if (rc) return rc; }
    if (status & kStatPosPoolFull) { int rc = growU32(reinterpret_cast<void**>(&m->dPosPool), m->posCap, sl.hStage->ctl[2], 4); if (rc) return rc; }
    if (status & kStatScratchFull) {
      // worst case of a strand: pmax intervals of < maxInterval entries, padded to a power of two
      uint64_t worst = 1;
      while (worst < static_cast<uint64_t>(m->pmax) * 1000ull * 2ull) worst <<= 1;
      if (m->scratchEntries >= worst) return fail(RAPMAP_ERR_CAPACITY, "hit-resolution work strip overflow at worst-case size");
      uint64_t ne = std::min<uint64_t>(worst, static_cast<uint64_t>(m->scratchEntries) * 8);
      cudaFree(m->dScratch); m->dScratch = nullptr;
      m->scratchEntries = static_cast<uint32_t>(ne);
      m->scratchStride = workAreaBytes(m->scratchEntries);
      // fewer resident warps when the strips get large
      while (m->gridMap > m->numSMs && m->scratchStride * static_cast<uint64_t>(m->gridMap) * kWarps > (8ull << 30)) m->gridMap -= m->numSMs;
      CU_TRY(cudaMalloc(&m->dScratch, m->scratchStride * static_cast<uint64_t>(m->gridMap) * kWarps));
    }
    if (status == 0 && hitsFull) {  // the merge produced more records than the hit arrays hold
      m->hitsCap = sl.hStage->mergeTotal + sl.hStage->mergeTotal / 4 + 1024;
      for (auto& any : m->slots) {
        cudaFree(any.dHits); any.dHits = nullptr;
        CU_TRY(cudaMalloc(&any.dHits, m->hitsCap * sizeof(rapmap_hit_t)));
        if (m->dopts.selAln) {
          cudaFree(any.dSelOut); any.dSelOut = nullptr;
          CU_TRY(cudaMalloc(&any.dSelOut, m->hitsCap * sizeof(rapmap_hit_t)));
        }
      }
      if (m->dopts.selAln) {
        cudaError_t e2 = selAlnReserve(m->selaln, m->hitsCap);
        if (e2 != cudaSuccess) return fail(RAPMAP_ERR_CUDA, std::string("selAlnReserve: ") + cudaGetErrorString(e2));
      }
    }
    int rc = enqueueAttempt(m, sl);
    if (rc) return rc;
  }
  total = m->dopts.selAln ? sl.hStage->selTotal : sl.hStage->mergeTotal;
  m->hitsPerPair = static_cast<double>(total) / static_cast<double>(n);
  // ---- results out
  out->num_hits = total;
  int rcOut = RAPMAP_OK;
  if (total > out->hits_capacity || (total > 0 && !out->hits)) {
    rcOut = fail(RAPMAP_ERR_CAPACITY, "hits_capacity too small for this batch (see num_hits)");
  } else if (!sl.directOut) {  // pageable host buffers: the size is known now, the copy takes a second synchronisation
    cudaMemcpyKind kind = out->location == RAPMAP_LOC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const rapmap_hit_t* src = m->dopts.selAln ? sl.dSelOut : sl.dHits;
    if (total > 0) CU_TRY(cudaMemcpyAsync(out->hits, src, total * sizeof(rapmap_hit_t), kind, m->sOut));
    CU_TRY(cudaMemcpyAsync(out->pair_offsets, sl.dPairOff, (n + 1) * 8, kind, m->sOut));
    CU_TRY(cudaEventRecord(sl.ev[7], m->sOut));
    CU_TRY(cudaEventRecord(sl.evOut, m->sOut));
    CU_TRY(waitSlot(m, sl));
  }
  for (int c = 0; c < 5; ++c) out->counters[c] = sl.hStage->counters.v[c];
  // paired reads: totHits is taken after the score filter (reference src/RapMapSAMapper.cpp:702); unmated: before (:241-246)
  if (m->dopts.selAln && sl.paired) out->counters[3] = total;
  auto ms = [&](int a, int b) { float t = 0; if (cudaEventElapsedTime(&t, sl.ev[a], sl.ev[b]) != cudaSuccess) { cudaGetLastError(); t = 0; } return t; };
  if (retries == 0) {  // the events of a repeated attempt do not line up: stage times are reported for clean batches only
    m->timing.ms_h2d = ms(0, 1);
    m->timing.ms_pack_reads = ms(11, 8);
    m->timing.ms_sa_collect = ms(8, 2);
    m->timing.ms_hits_to_mappings = ms(2, 3);
    m->timing.ms_merge = ms(3, 5);
    m->timing.ms_sel_aln = ms(5, 6);
    m->timing.ms_ksw = m->dopts.selAln ? ms(9, 10) : 0.0f;
    m->timing.ms_d2h = ms(6, 7);
    m->timing.ms_total = ms(0, 7);
  }
  m->timing.launches = sl.launches;
  m->timing.retries = retries;
  m->timing.sa_intervals = sl.hStage->ctl[0];
  m->timing.dp_jobs = sl.hStage->dpJobs[0];
  m->timing.dp_jobs_general = sl.hStage->dpJobs[1];
  m->timing.dp_jobs_exact_lane = sl.hStage->dpJobs[3];
  return rcOut;
}

int rapmap_cuda_mapper_create(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, uint64_t max_batch, uint32_t max_read_len,
                              rapmap_cuda_mapper_t** out) {
  return guarded([&] { return mapperCreateImpl(idx, opts, max_batch, max_read_len, out); });
}

void rapmap_cuda_mapper_free(rapmap_cuda_mapper_t* m) {
  if (!m) return;
  cudaSetDevice(m->idx->device);
  cudaStreamSynchronize(m->sIn); cudaStreamSynchronize(m->stream); cudaStreamSynchronize(m->sOut);
  freeMapperBuffers(m);
  delete m;
}

uint32_t rapmap_cuda_max_in_flight(void) { return static_cast<uint32_t>(kDepth); }

int rapmap_cuda_map_batch_async(rapmap_cuda_mapper_t* m, const rapmap_read_batch_t* reads, rapmap_hit_batch_t* out) {
  return guarded([&] { return mapBatchAsyncImpl(m, reads, out); });
}

int rapmap_cuda_mapper_wait(rapmap_cuda_mapper_t* m) {
  return guarded([&] { return mapperWaitImpl(m); });
}

int rapmap_cuda_map_batch(rapmap_cuda_mapper_t* m, const rapmap_read_batch_t* reads, rapmap_hit_batch_t* out) {
  return guarded([&] {
    int rc = mapBatchAsyncImpl(m, reads, out);
    if (rc) return rc;
    return mapperWaitImpl(m);
  });
}

int rapmap_cuda_last_timing(const rapmap_cuda_mapper_t* m, rapmap_cuda_timing_t* t) {
  if (!m || !t) return fail(RAPMAP_ERR_ARG, "null argument");
  *t = m->timing;
  return RAPMAP_OK;
}

void* rapmap_cuda_mapper_stream(const rapmap_cuda_mapper_t* m) { return m ? static_cast<void*>(m->stream) : nullptr; }

static int debugIntervalsImpl(rapmap_cuda_mapper_t* m, uint64_t read_index, rapmap_sa_interval_t* out, uint32_t cap, uint32_t* n_fwd,
                              uint32_t* n_rc, uint8_t* found_hit) {
  if (!m || !n_fwd || !n_rc || !found_hit) return fail(RAPMAP_ERR_ARG, "null argument");
  if (m->submitted != m->collected) return fail(RAPMAP_ERR_ARG, "a batch is in flight");
  if (read_index >= m->lastReads) return fail(RAPMAP_ERR_ARG, "read index out of range of the last batch");
  CU_TRY(cudaSetDevice(m->idx->device));
  ReadSummary s;
  CU_TRY(cudaMemcpy(&s, m->dSumm + read_index, sizeof(s), cudaMemcpyDeviceToHost));
  *n_fwd = s.nFwd; *n_rc = s.nRc; *found_hit = s.found;
  uint32_t tot = static_cast<uint32_t>(s.nFwd) + s.nRc;
  if (tot > cap || (tot > 0 && !out)) return fail(RAPMAP_ERR_CAPACITY, "interval buffer too small");
  if (tot == 0) return RAPMAP_OK;
  std::vector<IntervalRec> tmp(tot);
  CU_TRY(cudaMemcpy(tmp.data(), m->dIvArena + s.ivOff, tot * sizeof(IntervalRec), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < tot; ++i) {
    std::memset(&out[i], 0, sizeof(out[i]));
    out[i].begin = tmp[i].begin; out[i].end = tmp[i].end; out[i].len = tmp[i].len; out[i].query_pos = tmp[i].qpos;
    out[i].query_rc = i >= s.nFwd ? 1 : 0;
  }
  return RAPMAP_OK;
}

int rapmap_cuda_debug_intervals(rapmap_cuda_mapper_t* m, uint64_t read_index, rapmap_sa_interval_t* out, uint32_t cap, uint32_t* n_fwd,
                                uint32_t* n_rc, uint8_t* found_hit) {
  return guarded([&] { return debugIntervalsImpl(m, read_index, out, cap, n_fwd, n_rc, found_hit); });
}

static int dupOut(const std::string& s, char** sam, uint64_t* len) {
  char* p = static_cast<char*>(std::malloc(s.size() + 1));
  if (!p) return fail(RAPMAP_ERR_ARG, "out of host memory");
  std::memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  *sam = p;
  *len = s.size();
  return RAPMAP_OK;
}

int rapmap_cuda_sam_header(const rapmap_cuda_index_t* idx, char** sam, uint64_t* sam_len) {
  if (!idx || !sam || !sam_len) return fail(RAPMAP_ERR_ARG, "null argument");
  if (idx->names.size() != idx->hdr.numTxp || idx->lens.size() != idx->hdr.numTxp) return fail(RAPMAP_ERR_ARG, "index has no transcript name table");
  return guarded([&] { return dupOut(samHeader(idx->names, idx->lens), sam, sam_len); });
}

// Formats pairs [first, last) of the chunk into `out` (one worker of rapmap_cuda_format_sam).
static void formatRange(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, const rapmap_read_batch_t* reads, const std::vector<const char*>& n1,
                        const std::vector<const char*>& n2, rapmap_hit_batch_t* hits, uint64_t first, uint64_t last, std::string& out) {
  const bool paired = reads->seq2 != nullptr;
  out.reserve((last - first) * (paired ? 700 : 350));
  for (uint64_t i = first; i < last; ++i) {
    const char* s1; const char* s2 = nullptr; size_t l1, l2 = 0;
    if (reads->off1) {
      s1 = reinterpret_cast<const char*>(reads->seq1) + reads->off1[i]; l1 = reads->off1[i + 1] - reads->off1[i];
      if (paired) { s2 = reinterpret_cast<const char*>(reads->seq2) + reads->off2[i]; l2 = reads->off2[i + 1] - reads->off2[i]; }
    } else {
      s1 = reinterpret_cast<const char*>(reads->seq1) + i * reads->fixed_len; l1 = reads->fixed_len;
      if (paired) { s2 = reinterpret_cast<const char*>(reads->seq2) + i * reads->fixed_len; l2 = reads->fixed_len; }
    }
    const uint64_t b = hits->pair_offsets[i], e = hits->pair_offsets[i + 1];
    if (paired) samPair(idx->names, idx->lens, opts->max_num_hits, n1[i], s1, l1, n2[i], s2, l2, hits->hits + b, e - b, out);
    else samSingle(idx->names, idx->lens, n1[i], s1, l1, hits->hits + b, e - b, out);
  }
}

static int formatSamImpl(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, const rapmap_read_batch_t* reads, const char* names1,
                         const char* names2, rapmap_hit_batch_t* hits, uint32_t threads, char** sam, uint64_t* sam_len) {
  if (!idx || !opts || !reads || !names1 || !hits || !sam || !sam_len) return fail(RAPMAP_ERR_ARG, "null argument");
  if (reads->location != RAPMAP_LOC_HOST || hits->location != RAPMAP_LOC_HOST) return fail(RAPMAP_ERR_ARG, "format_sam needs host buffers");
  const bool paired = reads->seq2 != nullptr;
  if (paired && !names2) return fail(RAPMAP_ERR_ARG, "names2 is null for paired reads");
  if (idx->names.size() != idx->hdr.numTxp) return fail(RAPMAP_ERR_ARG, "index has no transcript name table");
  const uint64_t n = reads->n;
  std::vector<const char*> n1(n), n2(paired ? n : 0);
  {
    const char* p = names1;
    for (uint64_t i = 0; i < n; ++i) { n1[i] = p; p += std::strlen(p) + 1; }
    if (paired) { p = names2; for (uint64_t i = 0; i < n; ++i) { n2[i] = p; p += std::strlen(p) + 1; } }
  }
  const uint32_t T = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(threads ? threads : 1, (n + 4095) / 4096)));
  std::vector<std::string> parts(T);
  if (T == 1) formatRange(idx, opts, reads, n1, n2, hits, 0, n, parts[0]);
  else {
    std::vector<std::thread> th;
    std::vector<int> failed(T, 0);
    for (uint32_t t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        try { formatRange(idx, opts, reads, n1, n2, hits, n * t / T, n * (t + 1) / T, parts[t]); } catch (...) { failed[t] = 1; }
      });
    for (auto& x : th) x.join();
    for (int f : failed) if (f) return fail(RAPMAP_ERR_IO, "out of host memory while formatting SAM");
  }
  uint64_t total = 0;
  for (const auto& p : parts) total += p.size();
  char* buf = static_cast<char*>(std::malloc(total + 1));
  if (!buf) return fail(RAPMAP_ERR_ARG, "out of host memory");
  uint64_t at = 0;
  for (const auto& p : parts) { std::memcpy(buf + at, p.data(), p.size()); at += p.size(); }
  buf[total] = 0;
  *sam = buf;
  *sam_len = total;
  return RAPMAP_OK;
}

int rapmap_cuda_format_sam(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, const rapmap_read_batch_t* reads, const char* names1,
                           const char* names2, rapmap_hit_batch_t* hits, char** sam, uint64_t* sam_len) {
  return guarded([&] { return formatSamImpl(idx, opts, reads, names1, names2, hits, 1, sam, sam_len); });
}

int rapmap_cuda_format_sam_mt(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, const rapmap_read_batch_t* reads, const char* names1,
                              const char* names2, rapmap_hit_batch_t* hits, uint32_t threads, char** sam, uint64_t* sam_len) {
  return guarded([&] { return formatSamImpl(idx, opts, reads, names1, names2, hits, threads, sam, sam_len); });
}

void rapmap_cuda_free(void* p) { std::free(p); }

void* rapmap_cuda_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); fail(RAPMAP_ERR_CUDA, "cudaMallocHost failed"); return nullptr; }
  return p;
}

void rapmap_cuda_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

} // extern "C"
