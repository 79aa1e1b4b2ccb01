"""rapmap_b200 — B200-native quasi-mapping engine (drop-in for the `rapmap quasimap` per-read hot path).

Python host mirror of the C-ABI in ``include/rapmap_cuda.h`` (ctypes, no torch types in the boundary).
The classes follow the reference's objects for this path:

* :class:`Index`   — ``RapMapSAIndex`` (reference include/RapMapSAIndex.hpp:46-83), loaded unchanged from a
  ``quasiindex`` directory and uploaded once to HBM.
* :class:`Mapper`  — the per-thread ``SACollector`` + ``SASearcher`` + ``hit_manager`` + mate-merge set-up of
  ``processReadsPairSA`` (reference src/RapMapSAMapper.cpp:385-455); :meth:`Mapper.map_batch` is the body
  of its read loop (:461-711) for a whole chunk.

There is no CPU fallback: if the CUDA library is missing or no device is present, calls raise.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAPMAP_B200_LIB") or os.path.join(_HERE, "_build", "librapmap_cuda.so")  # env override: A/B builds of the same ABI

OK, ERR_IO, ERR_CUDA, ERR_UNSUPPORTED, ERR_ARG, ERR_CAPACITY = range(6)
LOC_HOST, LOC_DEVICE = 0, 1


class RapMapCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rapmap_cuda error {code}: {msg}")
        self.code = code


class Opts(C.Structure):
    """rapmap_cuda_opts_t (MappingOpts, reference src/RapMapSAMapper.cpp:114-152)."""

    _fields_ = [
        ("max_num_hits", C.c_uint32),
        ("quasi_coverage", C.c_double),
        ("sensitive", C.c_uint8),
        ("strict_check", C.c_uint8),
        ("fuzzy", C.c_uint8),
        ("sel_aln", C.c_uint8),
        ("consensus_slack", C.c_float),
        ("min_score_fraction", C.c_double),
        ("match_score", C.c_int16),
        ("mismatch_penalty", C.c_int16),
        ("gap_open_penalty", C.c_int16),
        ("gap_extend_penalty", C.c_int16),
        ("dp_bandwidth", C.c_int32),
        ("hard_filter", C.c_uint8),
        ("alignment_policy", C.c_uint8),
        ("no_orphans", C.c_uint8),
        ("no_dovetail", C.c_uint8),
        ("max_mmp_extension", C.c_int32),
        ("recover_orphans", C.c_uint8),
    ]


class Hit(C.Structure):
    """rapmap_hit_t — one QuasiAlignment (reference include/RapMapUtils.hpp:399-502)."""

    _fields_ = [
        ("tid", C.c_uint32),
        ("pos", C.c_int32),
        ("mate_pos", C.c_int32),
        ("frag_len", C.c_uint32),
        ("read_len", C.c_uint16),
        ("mate_len", C.c_uint16),
        ("aln_score", C.c_int32),
        ("fwd", C.c_uint8),
        ("mate_fwd", C.c_uint8),
        ("mate_status", C.c_uint8),
        ("chain_status", C.c_uint8),
    ]


HIT_DTYPE = np.dtype(
    [
        ("tid", "<u4"), ("pos", "<i4"), ("mate_pos", "<i4"), ("frag_len", "<u4"), ("read_len", "<u2"), ("mate_len", "<u2"),
        ("aln_score", "<i4"), ("fwd", "u1"), ("mate_fwd", "u1"), ("mate_status", "u1"), ("chain_status", "u1"),
    ]
)
assert HIT_DTYPE.itemsize == C.sizeof(Hit) == 28


class ReadBatch(C.Structure):
    _fields_ = [
        ("seq1", C.c_void_p), ("off1", C.c_void_p), ("seq2", C.c_void_p), ("off2", C.c_void_p),
        ("n", C.c_uint64), ("fixed_len", C.c_uint32), ("location", C.c_int32),
    ]


class HitBatch(C.Structure):
    _fields_ = [
        ("hits", C.c_void_p), ("hits_capacity", C.c_uint64), ("pair_offsets", C.c_void_p), ("num_hits", C.c_uint64),
        ("counters", C.c_uint64 * 5), ("location", C.c_int32),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("ms_h2d", C.c_float), ("ms_sa_collect", C.c_float), ("ms_hits_to_mappings", C.c_float), ("ms_merge", C.c_float),
        ("ms_sel_aln", C.c_float), ("ms_pack_reads", C.c_float), ("ms_d2h", C.c_float), ("ms_total", C.c_float),
        ("launches", C.c_uint32), ("retries", C.c_uint32), ("sa_intervals", C.c_uint64),
        ("ms_ksw", C.c_float), ("dp_jobs", C.c_uint32), ("dp_jobs_general", C.c_uint32), ("dp_jobs_exact_lane", C.c_uint32),
    ]


class SAInterval(C.Structure):
    _fields_ = [("begin", C.c_int64), ("end", C.c_int64), ("len", C.c_uint32), ("query_pos", C.c_uint32), ("query_rc", C.c_uint8), ("pad", C.c_uint8 * 7)]


# every symbol include/rapmap_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "rapmap_cuda_last_error", "rapmap_cuda_opts_default", "rapmap_cuda_opts_selaln", "rapmap_cuda_index_load", "rapmap_cuda_index_free",
    "rapmap_cuda_index_num_transcripts", "rapmap_cuda_index_transcript_name", "rapmap_cuda_index_transcript_len", "rapmap_cuda_index_k",
    "rapmap_cuda_index_device_bytes", "rapmap_cuda_index_image_bytes", "rapmap_cuda_index_image_ptr", "rapmap_cuda_index_from_image",
    "rapmap_cuda_mapper_create", "rapmap_cuda_mapper_free", "rapmap_cuda_map_batch", "rapmap_cuda_map_batch_async", "rapmap_cuda_mapper_wait", "rapmap_cuda_max_in_flight",
    "rapmap_cuda_last_timing", "rapmap_cuda_mapper_stream", "rapmap_cuda_debug_intervals",
    "rapmap_cuda_format_sam", "rapmap_cuda_format_sam_mt", "rapmap_cuda_sam_header", "rapmap_cuda_free", "rapmap_cuda_host_alloc", "rapmap_cuda_host_free",
]

_lib = None


def lib() -> C.CDLL:
    """Loads librapmap_cuda.so (built in-tree by ``__graft_entry__.build()``); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RapMapCudaError(ERR_CUDA, f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.rapmap_cuda_last_error.restype = C.c_char_p
    L.rapmap_cuda_opts_default.argtypes = [C.POINTER(Opts)]
    L.rapmap_cuda_opts_selaln.argtypes = [C.POINTER(Opts)]
    L.rapmap_cuda_index_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.rapmap_cuda_index_free.argtypes = [C.c_void_p]
    L.rapmap_cuda_index_num_transcripts.argtypes = [C.c_void_p]
    L.rapmap_cuda_index_num_transcripts.restype = C.c_uint64
    L.rapmap_cuda_index_transcript_name.argtypes = [C.c_void_p, C.c_uint64]
    L.rapmap_cuda_index_transcript_name.restype = C.c_char_p
    L.rapmap_cuda_index_transcript_len.argtypes = [C.c_void_p, C.c_uint64]
    L.rapmap_cuda_index_transcript_len.restype = C.c_uint64
    L.rapmap_cuda_index_k.argtypes = [C.c_void_p]
    L.rapmap_cuda_index_k.restype = C.c_uint32
    L.rapmap_cuda_index_device_bytes.argtypes = [C.c_void_p]
    L.rapmap_cuda_index_device_bytes.restype = C.c_uint64
    L.rapmap_cuda_index_image_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.rapmap_cuda_index_image_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.rapmap_cuda_index_from_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.rapmap_cuda_mapper_create.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p)]
    L.rapmap_cuda_mapper_free.argtypes = [C.c_void_p]
    L.rapmap_cuda_map_batch.argtypes = [C.c_void_p, C.POINTER(ReadBatch), C.POINTER(HitBatch)]
    L.rapmap_cuda_map_batch_async.argtypes = [C.c_void_p, C.POINTER(ReadBatch), C.POINTER(HitBatch)]
    L.rapmap_cuda_mapper_wait.argtypes = [C.c_void_p]
    L.rapmap_cuda_max_in_flight.restype = C.c_uint32
    L.rapmap_cuda_last_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    L.rapmap_cuda_mapper_stream.argtypes = [C.c_void_p]
    L.rapmap_cuda_mapper_stream.restype = C.c_void_p
    L.rapmap_cuda_debug_intervals.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(SAInterval), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    L.rapmap_cuda_format_sam.argtypes = [C.c_void_p, C.POINTER(Opts), C.POINTER(ReadBatch), C.c_char_p, C.c_char_p, C.POINTER(HitBatch), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.rapmap_cuda_format_sam_mt.argtypes = [C.c_void_p, C.POINTER(Opts), C.POINTER(ReadBatch), C.c_char_p, C.c_char_p, C.POINTER(HitBatch), C.c_uint32,
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.rapmap_cuda_sam_header.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.rapmap_cuda_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def max_in_flight() -> int:
    """Chunks a mapper pipelines (rapmap_cuda_max_in_flight)."""
    return int(lib().rapmap_cuda_max_in_flight())


def _check(rc: int) -> None:
    if rc != OK:
        raise RapMapCudaError(rc, lib().rapmap_cuda_last_error().decode("utf-8", "replace"))


def default_opts(sel_aln: bool = False, **kw) -> Opts:
    """CLI defaults of `rapmap quasimap` (reference src/RapMapSAMapper.cpp:992-1023); keyword overrides by field name."""
    o = Opts()
    (lib().rapmap_cuda_opts_selaln if sel_aln else lib().rapmap_cuda_opts_default)(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class Index:
    """RapMapSAIndex on one GPU. ``Index(dir, device)`` == ``rmi.load(dir)`` + upload."""

    def __init__(self, index_dir: Optional[str] = None, device: int = 0, _handle=None):
        self._h = C.c_void_p()
        self.device = device
        if _handle is not None:
            self._h = _handle
        else:
            _check(lib().rapmap_cuda_index_load(os.fsencode(index_dir), device, C.byref(self._h)))

    @classmethod
    def from_image(cls, device: int, device_ptr: int, nbytes: int, meta: Optional["Index"] = None) -> "Index":
        h = C.c_void_p()
        _check(lib().rapmap_cuda_index_from_image(meta._h if meta is not None else None, device, C.c_void_p(device_ptr), nbytes, C.byref(h)))
        return cls(device=device, _handle=h)

    def image(self):
        """(device pointer, bytes) of the packed device image — one broadcast replicates the index."""
        p, n = C.c_void_p(), C.c_uint64()
        _check(lib().rapmap_cuda_index_image_ptr(self._h, C.byref(p)))
        _check(lib().rapmap_cuda_index_image_bytes(self._h, C.byref(n)))
        return p.value, n.value

    @property
    def k(self) -> int:
        return lib().rapmap_cuda_index_k(self._h)

    @property
    def num_transcripts(self) -> int:
        return lib().rapmap_cuda_index_num_transcripts(self._h)

    @property
    def device_bytes(self) -> int:
        return lib().rapmap_cuda_index_device_bytes(self._h)

    def transcript_name(self, tid: int) -> str:
        return lib().rapmap_cuda_index_transcript_name(self._h, tid).decode()

    def transcript_len(self, tid: int) -> int:
        return lib().rapmap_cuda_index_transcript_len(self._h, tid)

    def sam_header(self) -> bytes:
        p, n = C.c_void_p(), C.c_uint64()
        _check(lib().rapmap_cuda_sam_header(self._h, C.byref(p), C.byref(n)))
        try:
            return C.string_at(p, n.value)
        finally:
            lib().rapmap_cuda_free(p)

    def close(self):
        if self._h:
            lib().rapmap_cuda_index_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class BatchResult:
    hits: np.ndarray          # HIT_DTYPE records, input order
    pair_offsets: np.ndarray  # uint64[n+1]
    counters: np.ndarray      # numReads, peHits, seHits, totHits, tooManyHits
    num_hits: int


def _as_ptr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())  # torch tensor


class Mapper:
    """One mapper per host thread / CUDA stream (like the reference's per-thread SACollector et al.)."""

    def __init__(self, index: Index, opts: Optional[Opts] = None, max_batch: int = 1 << 20, max_read_len: int = 100):
        self.index = index
        self.opts = opts if opts is not None else default_opts()
        self.max_batch = max_batch
        self.max_read_len = max_read_len
        self._h = C.c_void_p()
        _check(lib().rapmap_cuda_mapper_create(index._h, C.byref(self.opts), max_batch, max_read_len, C.byref(self._h)))
        self._hits_buf = None
        self._off_buf = None
        self._pending = collections.deque()

    def _read_batch(self, seq1, seq2, n, fixed_len, off1, off2, location) -> ReadBatch:
        rb = ReadBatch()
        rb.seq1 = _as_ptr(seq1)
        rb.seq2 = _as_ptr(seq2)
        rb.off1 = _as_ptr(off1)
        rb.off2 = _as_ptr(off2)
        rb.n = n
        rb.fixed_len = fixed_len
        rb.location = location
        return rb

    def map_batch(self, seq1, seq2=None, n: Optional[int] = None, fixed_len: int = 0, off1=None, off2=None,
                  location: int = LOC_HOST, hits_out=None, offsets_out=None, out_location: int = LOC_HOST, capacity: Optional[int] = None) -> BatchResult:
        """Maps one chunk. ``seq*``: uint8 numpy arrays (host) or torch CUDA tensors (location=LOC_DEVICE).
        Fixed-length reads: row-major ``n x fixed_len``; otherwise pass uint64 offset arrays of n+1 entries.
        Without ``hits_out`` the result arrays are views of buffers the mapper reuses: copy them before the next call."""
        if self._pending:
            raise RapMapCudaError(ERR_ARG, "map_batch with batches in flight: collect them with wait() first")
        self.map_batch_async(seq1, seq2, n, fixed_len, off1, off2, location, hits_out, offsets_out, out_location, capacity)
        try:
            return self.wait()
        except RapMapCudaError as e:
            if e.code == ERR_CAPACITY and hits_out is None:
                self._hits_buf = np.empty(int(self._last_hb.num_hits) + 1024, dtype=HIT_DTYPE)
                return self.map_batch(seq1, seq2, n, fixed_len, off1, off2, location, None, None, out_location, None)
            raise

    def map_batch_async(self, seq1, seq2=None, n: Optional[int] = None, fixed_len: int = 0, off1=None, off2=None,
                        location: int = LOC_HOST, hits_out=None, offsets_out=None, out_location: int = LOC_HOST, capacity: Optional[int] = None) -> None:
        """rapmap_cuda_map_batch_async: enqueues the chunk and returns; :meth:`wait` collects the OLDEST chunk in flight (a
        mapper keeps up to max_in_flight(): one computes while the next one's reads come in and the previous one's results go out).
        The input and output buffers must stay alive and untouched until then (the mapper keeps references); with chunks in
        flight pass your own ``hits_out`` / ``offsets_out`` per chunk."""
        if n is None:
            n = (len(off1) - 1) if off1 is not None else int(np.prod(seq1.shape)) // fixed_len
        rb = self._read_batch(seq1, seq2, n, fixed_len, off1, off2, location)
        hb = HitBatch()
        hb.location = out_location
        if hits_out is None:
            if self._pending:
                raise RapMapCudaError(ERR_ARG, "chunks in flight need caller-provided output buffers")
            cap = capacity if capacity is not None else max(1024, 8 * n)
            if self._hits_buf is None or len(self._hits_buf) < cap:
                self._hits_buf = np.empty(cap, dtype=HIT_DTYPE)
            if self._off_buf is None or len(self._off_buf) < n + 1:
                self._off_buf = np.empty(n + 1, dtype=np.uint64)
            hits_out, offsets_out = self._hits_buf, self._off_buf
        hb.hits = _as_ptr(hits_out)
        hb.hits_capacity = capacity if capacity is not None else (len(hits_out) if isinstance(hits_out, np.ndarray) else hits_out.numel() // 28)
        hb.pair_offsets = _as_ptr(offsets_out)
        _check(lib().rapmap_cuda_map_batch_async(self._h, C.byref(rb), C.byref(hb)))
        self._pending.append((rb, hb, hits_out, offsets_out, n, (seq1, seq2, off1, off2)))

    @property
    def in_flight(self) -> int:
        return len(self._pending)

    def wait(self) -> BatchResult:
        """rapmap_cuda_mapper_wait: blocks until the oldest chunk in flight is mapped and its result is in its output buffers."""
        if not self._pending:
            raise RapMapCudaError(ERR_ARG, "no batch in flight")
        rb, hb, hits_out, offsets_out, n, _keep = self._pending.popleft()
        self._last_hb = hb
        _check(lib().rapmap_cuda_mapper_wait(self._h))
        nh = int(hb.num_hits)
        if isinstance(hits_out, np.ndarray):
            return BatchResult(hits_out[:nh], offsets_out[: n + 1], np.array(list(hb.counters), dtype=np.uint64), nh)
        return BatchResult(hits_out, offsets_out, np.array(list(hb.counters), dtype=np.uint64), nh)

    def timing(self) -> Timing:
        t = Timing()
        _check(lib().rapmap_cuda_last_timing(self._h, C.byref(t)))
        return t

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of this mapper (all of map_batch is issued on it)."""
        return int(lib().rapmap_cuda_mapper_stream(self._h) or 0)

    def debug_intervals(self, read_index: int, cap: int = 2048):
        """SAIntervalHit lists of read `read_index` of the last batch (mate-1 reads first, then mate-2)."""
        buf = (SAInterval * cap)()
        nf, nr, found = C.c_uint32(), C.c_uint32(), C.c_uint8()
        _check(lib().rapmap_cuda_debug_intervals(self._h, read_index, buf, cap, C.byref(nf), C.byref(nr), C.byref(found)))
        ivs = [(int(b.begin), int(b.end), int(b.len), int(b.query_pos), int(b.query_rc)) for b in buf[: nf.value + nr.value]]
        return bool(found.value), nf.value, nr.value, ivs

    def format_sam(self, seq1, seq2, names1, names2, result: BatchResult, n: int, fixed_len: int = 0, off1=None, off2=None, threads: int = 1) -> bytes:
        """SAM text of a chunk (host side; reference src/RapMapUtils.cpp:230-588). names*: lists of str; seq2 / names2 None for
        unmated reads."""
        rb = self._read_batch(seq1, seq2, n, fixed_len, off1, off2, LOC_HOST)
        hb = HitBatch()
        hits = np.ascontiguousarray(result.hits)
        offs = np.ascontiguousarray(result.pair_offsets)
        hb.hits = hits.ctypes.data
        hb.hits_capacity = len(hits)
        hb.pair_offsets = offs.ctypes.data
        hb.num_hits = result.num_hits
        n1 = b"\0".join(s.encode() for s in names1) + b"\0"
        n2 = (b"\0".join(s.encode() for s in names2) + b"\0") if names2 is not None else None
        p, ln = C.c_void_p(), C.c_uint64()
        _check(lib().rapmap_cuda_format_sam_mt(self.index._h, C.byref(self.opts), C.byref(rb), n1, n2, C.byref(hb), threads, C.byref(p), C.byref(ln)))
        try:
            return C.string_at(p, ln.value)
        finally:
            lib().rapmap_cuda_free(p)

    def close(self):
        if self._h:
            lib().rapmap_cuda_mapper_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
