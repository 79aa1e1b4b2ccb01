// Kernel-side data structures shared between kernels.cu and capi.cu.
#pragma once
#include <cstdint>
#include "device_index.cuh"
#include "../../include/rapmap_cuda.h"

namespace rapmap_b200 {

// A chunk of reads as the kernels see it (device pointers).  Read r in [0, numReads): mate = r >= n.
struct BatchView {
  const uint8_t* seq[2];
  const uint64_t* off[2];
  uint32_t fixedLen;
  uint64_t n;         // pairs (or unmated reads)
  uint64_t numReads;  // n or 2n
};

// One SAIntervalHit (reference include/RapMapUtils.hpp:516-525); queryRC is implied by the list it is in.
struct IntervalRec {
  int32_t begin, end;
  uint16_t len, qpos;
};

// Per-read result of the SA-lookup kernel (HitCollectorInfo, reference include/HitManager.hpp:59-72).
struct ReadSummary {
  uint32_t ivOff;        // first record in the interval arena: nFwd forward records, then nRc
  uint16_t nFwd, nRc;
  uint16_t readLen;
  uint8_t found;         // return value of SACollector::operator()
  uint8_t pad;
};

// One QuasiAlignment of a single read before mate merging (reference src/HitManager.cpp:691-882).
struct QARec {
  uint32_t tid;
  int32_t pos;
  uint32_t posOff;       // allPositions in the position pool (only when fuzzy / selAln)
  uint32_t nAll;
  uint32_t oppOff;       // oppositeStrandPositions
  uint32_t nOpp;
  uint8_t fwd;
  uint8_t chain;         // ChainStatus of this read end
  uint16_t pad;
};

struct QASummary {
  uint32_t qaOff;
  uint32_t nQA;
};

// Mapping options in the form the kernels consume (derived as reference src/RapMapSAMapper.cpp:385-455).
struct DevOpts {
  uint32_t maxNumHits;
  double covReq;
  float consensusFraction;
  int32_t maxMMPExtension;
  int32_t strictCheckSlack;
  int32_t maxInterval;
  uint8_t doChaining, considerMultiPos, fuzzy, selAln;
  uint8_t disableNIP, strictCheck;   // SACollector::disableNIP_ / strictCheck_ (coverage mode when both are set)
  uint8_t noOrphans, noDovetail, hardFilter, alignmentPolicy;
  uint8_t recoverOrphans;            // --recoverOrphans (acts only with the fuzzy merge: -s / -f)
  int16_t ma, mm, go, ge;
  int32_t dpBandwidth;
  double minScoreFraction;
};

struct Counters5 { unsigned long long v[5]; };  // numReads, peHits, seHits, totHits, tooManyHits

// Status word written by the kernels (host checks it after the batch).
enum : uint32_t {
  kStatIntervalArenaFull = 1u,
  kStatQAArenaFull = 2u,
  kStatScratchFull = 4u,
  kStatReadTooLong = 8u,
  kStatPosPoolFull = 16u,
  kStatIvScratchFull = 32u,
  kStatInternal = 64u,        // a kernel met a case its launch configuration excludes (a bug, never a capacity matter)
};

} // namespace rapmap_b200
