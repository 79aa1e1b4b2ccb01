/*
 * rapmap_cuda.h — C-ABI of the B200 quasi-mapping engine (librapmap_cuda.so).
 *
 * The reference (COMBINE-lab/RapMap v0.6.0) has no FFI / plugin layer; its boundary for this path is
 * the per-chunk body of processReadsPairSA (src/RapMapSAMapper.cpp:461-711): a ReadGroup chunk of read
 * pairs goes in, a vector<QuasiAlignment> per pair comes out.  These entry points replace exactly that
 * body, one call per chunk ("batch"), and nothing else.  All paths below are relative to the reference
 * root.  Plain C types only; no exceptions or exit() cross this boundary; every function returns an
 * int status (0 = RAPMAP_OK) unless stated otherwise.
 *
 * There is NO CPU fallback behind these symbols: without a CUDA device (or for an option the device
 * path does not implement) the calls fail with an error code and a message in rapmap_cuda_last_error().
 */
#ifndef RAPMAP_CUDA_H
#define RAPMAP_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  RAPMAP_OK = 0,
  RAPMAP_ERR_IO = 1,           /* index files missing / malformed                                */
  RAPMAP_ERR_CUDA = 2,         /* CUDA runtime error (no device, OOM, launch failure)            */
  RAPMAP_ERR_UNSUPPORTED = 3,  /* option or index flavour not implemented on the device path     */
  RAPMAP_ERR_ARG = 4,          /* bad argument                                                   */
  RAPMAP_ERR_CAPACITY = 5      /* caller-provided output buffer too small (num_hits says how big) */
};

typedef struct rapmap_cuda_index rapmap_cuda_index_t;   /* replaces RapMapSAIndex<IndexT,HashT>, include/RapMapSAIndex.hpp:46-83 */
typedef struct rapmap_cuda_mapper rapmap_cuda_mapper_t; /* replaces the per-thread SACollector + SASearcher + KSW2Aligner + AlnCacheMap
                                                           set-up of processReadsPairSA, src/RapMapSAMapper.cpp:385-455 */

/* POD mirror of the MappingOpts fields that affect mapping (src/RapMapSAMapper.cpp:114-152) with the
 * CLI defaults of :992-1023.  rapmap_cuda_opts_default() fills the defaults of plain `quasimap`;
 * rapmap_cuda_opts_selaln() those of `quasimap -s`. */
typedef struct {
  uint32_t max_num_hits;        /* -m / --maxNumHits            (200)  */
  double quasi_coverage;        /* -z / --quasiCoverage         (0.0)  */
  uint8_t sensitive;            /* !--noSensitive               (1)  => NIP skipping off, SACollector::disableNIP  */
  uint8_t strict_check;         /* !--noStrictCheck             (1)    */
  uint8_t fuzzy;                /* -f / --fuzzyIntersection     (0)    */
  uint8_t sel_aln;              /* -s / --selAln                (0)  => chaining, multi-position, fuzzy merge, ksw2 */
  float consensus_slack;        /* --consensusSlack             (0.2, read only with sel_aln) */
  double min_score_fraction;    /* --minScoreFrac               (0.65) */
  int16_t match_score;          /* --ma (2)  */
  int16_t mismatch_penalty;     /* --mm (-4) */
  int16_t gap_open_penalty;     /* --go (4)  */
  int16_t gap_extend_penalty;   /* --ge (2)  */
  int32_t dp_bandwidth;         /* --dpBandwidth (15) */
  uint8_t hard_filter;          /* --hardFilter */
  uint8_t alignment_policy;     /* 0 DEFAULT, 1 BT2 (--mimicBT2), 2 BT2_STRICT (--mimicStrictBT2); include/SelectiveAlignmentUtils.hpp:23-27 */
  uint8_t no_orphans;           /* --noOrphans  */
  uint8_t no_dovetail;          /* --noDovetail */
  int32_t max_mmp_extension;    /* --maxMMPExtension (7) */
  uint8_t recover_orphans;      /* --recoverOrphans: orphan rescue next to the anchor hit (include/SelectiveAlignmentUtils.hpp:35-257); acts with sel_aln / fuzzy */
} rapmap_cuda_opts_t;

/* One QuasiAlignment (include/RapMapUtils.hpp:399-502), the fields that reach SAM / Salmon. 28 bytes. */
typedef struct {
  uint32_t tid;          /* transcript id                                                      */
  int32_t pos;           /* leftmost position of this read on the transcript (may be < 0)      */
  int32_t mate_pos;      /* valid iff mate_status == 3 (PAIRED_END_PAIRED), else 0             */
  uint32_t frag_len;     /* 0 for orphans / unmated                                            */
  uint16_t read_len;
  uint16_t mate_len;     /* valid iff paired, else 0                                           */
  int32_t aln_score;     /* QuasiAlignment::alnScore_ (0 unless sel_aln)                       */
  uint8_t fwd;           /* read maps to the forward strand                                    */
  uint8_t mate_fwd;      /* mateIsFwd (1 for orphans, reference sets it true)                  */
  uint8_t mate_status;   /* MateStatus: 0 SINGLE_END, 1 PAIRED_END_LEFT, 2 PAIRED_END_RIGHT, 3 PAIRED_END_PAIRED */
  uint8_t chain_status;  /* FragmentChainStatus: left | right << 4 (ChainStatus 0 PERFECT 1 UNGAPPED 4 REGULAR)  */
} rapmap_hit_t;

#define RAPMAP_LOC_HOST 0
#define RAPMAP_LOC_DEVICE 1

/* A chunk of reads (fastx_parser::ReadGroup of ReadPair, include/FastxParser.hpp:62-71): concatenated
 * ASCII bases.  If off1/off2 are NULL every read has exactly fixed_len bases (row-major n x fixed_len);
 * otherwise read i of mate m is seq_m[off_m[i] .. off_m[i+1]).  seq2 == NULL means unmated reads
 * (processReadsSingleSA, src/RapMapSAMapper.cpp:156-371).  location says whether the pointers are host
 * or device (cudaMalloc) memory. */
typedef struct {
  const uint8_t* seq1;
  const uint64_t* off1;
  const uint8_t* seq2;
  const uint64_t* off2;
  uint64_t n;
  uint32_t fixed_len;
  int32_t location;
} rapmap_read_batch_t;

/* Result of one chunk, in input order: hits[pair_offsets[i] .. pair_offsets[i+1]) are pair i's jointHits
 * in the reference's order.  hits / pair_offsets are caller-provided (host, ideally pinned, or device
 * per `location`); pair_offsets must hold n+1 entries.  counters mirrors HitCounters
 * (include/RapMapUtils.hpp:208-216): numReads, peHits, seHits, totHits, tooManyHits. */
typedef struct {
  rapmap_hit_t* hits;
  uint64_t hits_capacity;
  uint64_t* pair_offsets;
  uint64_t num_hits;
  uint64_t counters[5];
  int32_t location;
} rapmap_hit_batch_t;

/* Per-stage device time of the last rapmap_cuda_map_batch call, measured with CUDA events on the
 * mapper's stream (milliseconds), and launch counts. */
typedef struct {
  float ms_h2d, ms_sa_collect, ms_hits_to_mappings, ms_merge, ms_sel_aln, ms_pack_reads, ms_d2h, ms_total;  /* ms_sa_collect: the SA-lookup kernel alone; ms_pack_reads: the 2-bit read packer before it */
  uint32_t launches;        /* kernels launched by the call */
  uint32_t retries;         /* arena-overflow relaunches    */
  uint64_t sa_intervals;    /* SAIntervalHit records produced by the SA-lookup kernel */
  float ms_ksw;             /* the ksw2 DP kernels alone (part of ms_sel_aln) */
  uint32_t dp_jobs;         /* ksw_extz2 DP problems of the batch (getAlnScore calls that reach the aligner) */
  uint32_t dp_jobs_general; /* of those, the ones the thread-per-job kernel left to the general warp-per-job kernel */
  uint32_t dp_jobs_exact_lane; /* of those, the ones the two-jobs-per-thread kernel handed to the byte-exact thread-per-job kernel */
} rapmap_cuda_timing_t;

/* Thread-local message for the last non-zero status. */
const char* rapmap_cuda_last_error(void);

void rapmap_cuda_opts_default(rapmap_cuda_opts_t* o);
void rapmap_cuda_opts_selaln(rapmap_cuda_opts_t* o);

/* RapMapSAIndex::load (src/RapMapSAIndex.cpp:96-176): parses header.json, sa.bin, txpInfo.bin, rsd.bin and
 * hash.bin (dense) or hash_info.bph/.val (-p) unchanged, builds the device image and uploads it once to
 * the HBM of `device`. */
int rapmap_cuda_index_load(const char* index_dir, int device, rapmap_cuda_index_t** out);
void rapmap_cuda_index_free(rapmap_cuda_index_t* idx);

/* Index metadata (host side; what writeSAMHeader / the SAM formatter need, include/RapMapUtils.hpp:95-110). */
uint64_t rapmap_cuda_index_num_transcripts(const rapmap_cuda_index_t* idx);
const char* rapmap_cuda_index_transcript_name(const rapmap_cuda_index_t* idx, uint64_t tid);
uint64_t rapmap_cuda_index_transcript_len(const rapmap_cuda_index_t* idx, uint64_t tid);
uint32_t rapmap_cuda_index_k(const rapmap_cuda_index_t* idx);
uint64_t rapmap_cuda_index_device_bytes(const rapmap_cuda_index_t* idx);
/* Packed device image, for replicating the index to other GPUs (one ncclBroadcast of this blob):
 * export copies the image to a caller buffer of `bytes` (device memory on idx's device);
 * import builds an index object on `device` around a received blob (256-byte aligned device memory that the caller
 * keeps alive); the image carries the transcript names and lengths, so meta_src may be NULL. */
int rapmap_cuda_index_image_bytes(const rapmap_cuda_index_t* idx, uint64_t* bytes);
int rapmap_cuda_index_image_ptr(const rapmap_cuda_index_t* idx, void** device_ptr);
int rapmap_cuda_index_from_image(const rapmap_cuda_index_t* meta_src, int device, void* device_blob, uint64_t bytes,
                                 rapmap_cuda_index_t** out);

/* One mapper per host thread / CUDA stream (the reference's per-thread objects are not thread-safe either,
 * SURVEY.md §8b).  max_batch = largest `n` that will be passed to map_batch (sizes the device work areas). */
int rapmap_cuda_mapper_create(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, uint64_t max_batch,
                              uint32_t max_read_len, rapmap_cuda_mapper_t** out);
void rapmap_cuda_mapper_free(rapmap_cuda_mapper_t* m);

/* The hot path: body of the `for (auto& rpair : rg)` loop of processReadsPairSA for a whole chunk
 * (src/RapMapSAMapper.cpp:461-711), up to and excluding SAM formatting. */
int rapmap_cuda_map_batch(rapmap_cuda_mapper_t* m, const rapmap_read_batch_t* reads, rapmap_hit_batch_t* out);
/* The same call split in two, so that ONE host thread keeps the device busy - the reference gets its overlap from N worker
 * threads around a shared parser queue (src/RapMapSAMapper.cpp:853-909); here a mapper pipelines consecutive chunks on three
 * CUDA streams: while chunk i computes, the reads of chunk i+1 come in over PCIe and the results of chunk i-1 go out.
 * _async validates, enqueues the copies and every kernel of the chunk and returns without waiting; up to
 * rapmap_cuda_max_in_flight() chunks (3) may be in flight per mapper (one more _async fails with RAPMAP_ERR_ARG).  `reads`, its buffers, `out` and its buffers must stay
 * valid and untouched until the rapmap_cuda_mapper_wait(m) that collects the chunk returns; _wait collects the OLDEST chunk
 * in flight and fills its out->num_hits / counters and the caller's buffers.  With pinned (or device) output buffers a chunk
 * costs one host synchronisation; pageable output buffers cost a second one.  rapmap_cuda_map_batch == _async + _wait. */
int rapmap_cuda_map_batch_async(rapmap_cuda_mapper_t* m, const rapmap_read_batch_t* reads, rapmap_hit_batch_t* out);
uint32_t rapmap_cuda_max_in_flight(void);
int rapmap_cuda_mapper_wait(rapmap_cuda_mapper_t* m);
int rapmap_cuda_last_timing(const rapmap_cuda_mapper_t* m, rapmap_cuda_timing_t* t);
/* The mapper's CUDA stream (cudaStream_t as void*): every kernel and copy of map_batch is issued on it, so a caller
 * can bracket calls with its own CUDA events. */
void* rapmap_cuda_mapper_stream(const rapmap_cuda_mapper_t* m);

/* Stage taps for parity tests: SAIntervalHit lists exactly as SACollector::operator() leaves them in
 * HitCollectorInfo (include/SACollector.hpp:108-362, include/HitManager.hpp:59-72) for the reads of the
 * LAST map_batch call (mate1 reads first, then mate2).  Host buffers. */
typedef struct {
  int64_t begin, end;      /* half-open SA interval */
  uint32_t len, query_pos;
  uint8_t query_rc;
  uint8_t pad[7];
} rapmap_sa_interval_t;
int rapmap_cuda_debug_intervals(rapmap_cuda_mapper_t* m, uint64_t read_index, rapmap_sa_interval_t* out, uint32_t cap,
                                uint32_t* n_fwd, uint32_t* n_rc, uint8_t* found_hit);

/* Host-side SAM text for a chunk: paired reads (writeAlignmentsToStream / writeUnalignedPairToStream,
 * src/RapMapUtils.cpp:313-588,137-196) or, when reads->seq2 == NULL, unmated reads (:198-311; names2 may be NULL);
 * include/RapMapUtils.hpp:687-810 for flags and soft clipping.  names are '\0'-separated per read.  Returns a malloc'ed
 * buffer the caller frees with rapmap_cuda_free.  _mt splits the chunk over `threads` host threads (the reference formats
 * inside its N worker threads); the text is identical. */
int rapmap_cuda_format_sam(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, const rapmap_read_batch_t* reads,
                           const char* names1, const char* names2, rapmap_hit_batch_t* hits, char** sam, uint64_t* sam_len);
int rapmap_cuda_format_sam_mt(const rapmap_cuda_index_t* idx, const rapmap_cuda_opts_t* opts, const rapmap_read_batch_t* reads,
                              const char* names1, const char* names2, rapmap_hit_batch_t* hits, uint32_t threads, char** sam, uint64_t* sam_len);
int rapmap_cuda_sam_header(const rapmap_cuda_index_t* idx, char** sam, uint64_t* sam_len);
void rapmap_cuda_free(void* p);

/* Page-locked host memory for read / result buffers (callers that do not link the CUDA runtime themselves).  With pinned
 * buffers the copies of rapmap_cuda_map_batch_async overlap the kernels and a chunk costs one host synchronisation. */
void* rapmap_cuda_host_alloc(uint64_t bytes);
void rapmap_cuda_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* RAPMAP_CUDA_H */
