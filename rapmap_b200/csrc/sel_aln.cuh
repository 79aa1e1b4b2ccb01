// Kernel 4 — selective alignment (getAlnScore + ksw_extz2_sse restatement + score filter).
// Placeholder until the device implementation lands: selecting -s fails loudly (no CPU fallback).
#pragma once
#include <string>
#include "kernels.cuh"

namespace rapmap_b200 {

struct SelAlnWork {
  rapmap_hit_t* outHits{nullptr};
};

inline cudaError_t selAlnAlloc(SelAlnWork&, uint64_t, uint32_t) { return cudaSuccess; }
inline void selAlnFree(SelAlnWork&) {}
inline int selAlnRun(SelAlnWork&, const DeviceIndex&, const DevOpts&, const BatchView&, uint64_t, bool, rapmap_hit_t*, uint64_t*, uint32_t*, uint64_t,
                     void*, size_t, int, cudaStream_t, uint32_t*, uint64_t*, void*, std::string& err) {
  err = "selective alignment is not implemented on the device path yet";
  return RAPMAP_ERR_UNSUPPORTED;
}

} // namespace rapmap_b200
