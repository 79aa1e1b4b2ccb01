"""Test-side plumbing: the CPU oracle (ctypes), the compiled reference binary, synthetic data, golden fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module; the product
package (rapmap_b200/) never touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import gzip
import hashlib
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

import rapmap_b200 as rb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CACHE = os.environ.get("RAPMAP_B200_CACHE", "/tmp/rapmap_b200_cache")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
ORACLE_CLI = os.path.join(ROOT, "oracle", "_build", "quasimap_oracle")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "rapmap_ref")
SYNTH_SO = os.path.join(ROOT, "build", "libsynth.so")
SYNTH_BIN = os.path.join(ROOT, "build", "bin", "synth")


# ----------------------------------------------------------------------------------------------
# oracle
# ----------------------------------------------------------------------------------------------
_olib = None
_oidx = {}


def oracle_lib() -> C.CDLL:
    global _olib
    if _olib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
        L = C.CDLL(ORACLE_SO)
        L.oracle_index_load.restype = C.c_void_p
        L.oracle_index_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.oracle_index_free.argtypes = [C.c_void_p]
        L.oracle_index_num_kmers.restype = C.c_uint64
        L.oracle_index_num_kmers.argtypes = [C.c_void_p]
        L.oracle_mapper_new.restype = C.c_void_p
        L.oracle_mapper_new.argtypes = [C.c_void_p, C.POINTER(rb.Opts)]
        L.oracle_mapper_free.argtypes = [C.c_void_p]
        L.oracle_map_batch.argtypes = [C.c_void_p, C.POINTER(rb.ReadBatch), C.POINTER(rb.HitBatch)]
        L.oracle_collect.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(rb.SAInterval), C.c_uint32, C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
        L.oracle_op_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_ksw_extz_score.restype = C.c_int32
        L.oracle_ksw_extz_score.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        _olib = L
    return _olib


def oracle_index(idx_dir: str):
    if idx_dir not in _oidx:
        err = C.create_string_buffer(512)
        h = oracle_lib().oracle_index_load(os.fsencode(idx_dir), err, 512)
        if not h:
            raise RuntimeError("oracle index load failed: " + err.value.decode())
        _oidx[idx_dir] = h
    return _oidx[idx_dir]


class OracleMapper:
    def __init__(self, idx_dir: str, opts: rb.Opts):
        self.L = oracle_lib()
        self.h = self.L.oracle_mapper_new(oracle_index(idx_dir), C.byref(opts))

    def map(self, s1, s2, fixed_len=0, off1=None, off2=None, n=None) -> rb.BatchResult:
        if n is None:
            n = (len(off1) - 1) if off1 is not None else s1.size // fixed_len
        rbatch = rb.ReadBatch()
        rbatch.seq1 = s1.ctypes.data
        rbatch.seq2 = s2.ctypes.data if s2 is not None else 0
        rbatch.off1 = off1.ctypes.data if off1 is not None else 0
        rbatch.off2 = off2.ctypes.data if off2 is not None else 0
        rbatch.n = n
        rbatch.fixed_len = fixed_len
        cap = max(1024, 16 * n)
        while True:
            hits = np.zeros(cap, dtype=rb.HIT_DTYPE)
            offs = np.zeros(n + 1, dtype=np.uint64)
            hb = rb.HitBatch()
            hb.hits = hits.ctypes.data
            hb.hits_capacity = cap
            hb.pair_offsets = offs.ctypes.data
            rc = self.L.oracle_map_batch(self.h, C.byref(rbatch), C.byref(hb))
            if rc == 5:
                cap = int(hb.num_hits) + 16
                continue
            assert rc == 0
            nh = int(hb.num_hits)
            return rb.BatchResult(hits[:nh], offs, np.array(list(hb.counters), dtype=np.uint64), nh)

    def collect(self, read: bytes, cap: int = 4096):
        buf = (rb.SAInterval * cap)()
        nf, nr, found = C.c_uint32(), C.c_uint32(), C.c_uint8()
        rc = self.L.oracle_collect(self.h, read, len(read), buf, cap, C.byref(nf), C.byref(nr), C.byref(found))
        assert rc == 0
        ivs = [(int(b.begin), int(b.end), int(b.len), int(b.query_pos), int(b.query_rc)) for b in buf[: nf.value + nr.value]]
        return bool(found.value), nf.value, nr.value, ivs

    def op_counts(self):
        out = (C.c_uint64 * 8)()
        self.L.oracle_op_counts(self.h, out)
        return dict(zip(["hashFind", "saProbes", "textCmp", "rankCalls", "intervals", "kswCalls", "alnCalls", "kswCells"], [int(x) for x in out]))

    def ksw(self, q: bytes, t: bytes) -> int:
        return self.L.oracle_ksw_extz_score(self.h, q, len(q), t, len(t))

    def __del__(self):
        try:
            self.L.oracle_mapper_free(self.h)
        except Exception:
            pass


def oracle_map(idx_dir, opts, s1, s2, fixed_len=0, off1=None, off2=None) -> rb.BatchResult:
    return OracleMapper(idx_dir, opts).map(s1, s2, fixed_len, off1, off2)


# ----------------------------------------------------------------------------------------------
# synthetic data (tools/synth.cpp)
# ----------------------------------------------------------------------------------------------
_slib = None


def synth_lib() -> C.CDLL:
    global _slib
    if _slib is None:
        if not os.path.exists(SYNTH_SO):
            os.makedirs(os.path.dirname(SYNTH_SO), exist_ok=True)
            subprocess.run(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", os.path.join(ROOT, "tools", "synth.cpp"), "-o", SYNTH_SO], check=True)
        L = C.CDLL(SYNTH_SO)
        L.synth_txome_new.restype = C.c_void_p
        L.synth_txome_new.argtypes = [C.c_uint64, C.c_int64, C.c_int]
        L.synth_txome_free.argtypes = [C.c_void_p]
        L.synth_txome_ntxp.restype = C.c_int64
        L.synth_txome_ntxp.argtypes = [C.c_void_p]
        L.synth_txome_text_len.restype = C.c_int64
        L.synth_txome_text_len.argtypes = [C.c_void_p]
        L.synth_txome_write_fasta.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_char_p]
        L.synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        _slib = L
    return _slib


class SynthTxome:
    def __init__(self, seed: int, genes: int, repeats: int = 0):
        self.seed, self.genes, self.repeats = seed, genes, repeats
        self.h = synth_lib().synth_txome_new(seed, genes, repeats)

    @property
    def ntxp(self):
        return synth_lib().synth_txome_ntxp(self.h)

    def write_fasta(self, path: str):
        assert synth_lib().synth_txome_write_fasta(self.seed, self.genes, self.repeats, os.fsencode(path)) == 0

    def reads(self, n: int, rseed: int = 54321, first: int = 0, read_len: int = 100, sub=10000, ins=300, dele=300, nn=1000, out1=None, out2=None):
        s1 = out1 if out1 is not None else np.empty((n, read_len), dtype=np.uint8)
        s2 = out2 if out2 is not None else np.empty((n, read_len), dtype=np.uint8)
        p1 = s1.ctypes.data if isinstance(s1, np.ndarray) else int(s1.data_ptr())
        p2 = s2.ctypes.data if isinstance(s2, np.ndarray) else int(s2.data_ptr())
        synth_lib().synth_reads(self.h, rseed, first, n, read_len, sub, ins, dele, nn, p1, p2, None)
        return s1, s2

    def __del__(self):
        try:
            synth_lib().synth_txome_free(self.h)
        except Exception:
            pass


def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def build_index(fasta: str, out_dir: str, perfect: bool = False, k: int = 31) -> str:
    """quasiindex with the compiled reference (oracle/_ref) — the index format is the reference's own."""
    if not out_dir.endswith("/"):
        out_dir += "/"
    if os.path.exists(os.path.join(out_dir, "header.json")):
        return out_dir
    os.makedirs(out_dir, exist_ok=True)
    cmd = [REF_BIN, "quasiindex", "-t", fasta, "-i", out_dir, "-k", str(k)]
    if perfect:
        cmd += ["-p", "-x", "4"]
    subprocess.run(cmd, check=True, capture_output=True)
    return out_dir


def synth_index(genes: int, seed: int = 12345, repeats: int = 0, perfect: bool = False) -> tuple[str, SynthTxome]:
    """Index of a synthetic transcriptome, cached under CACHE (built with the reference's quasiindex)."""
    tag = f"synth_g{genes}_s{seed}_r{repeats}" + ("_p" if perfect else "")
    d = os.path.join(CACHE, tag)
    tx = SynthTxome(seed, genes, repeats)
    if not os.path.exists(os.path.join(d, "idx", "header.json")):
        os.makedirs(d, exist_ok=True)
        fa = os.path.join(d, "t.fasta")
        tx.write_fasta(fa)
        build_index(fa, os.path.join(d, "idx"), perfect)
    return os.path.join(d, "idx") + "/", tx


ANTISENSE = {"seed": 2024, "genes": 60}


def antisense_index() -> tuple[str, SynthTxome, str]:
    """Index of a small synthetic transcriptome to which diverged ANTISENSE copies are added: for every second
    transcript, the reverse complement of bases [60, 460) with a substitution every 41-67 bases, between random flanks.
    A read of such a transcript has k-mer hits on BOTH strands with different coverage, so the strand decision of
    SACollector (coverage vs k-mer votes vs none: --noStrictCheck / --noSensitive) and the NIP skip change the result -
    on the plain synthetic transcriptomes those flags provably change nothing (tests/golden/golden.json: synth/nosensitive
    == synth/nostrict == synth/default).  Built with the reference's quasiindex, cached under CACHE.
    Returns (index dir, transcriptome the reads are drawn from, fasta path)."""
    import random

    d = os.path.join(CACHE, f"antisense_g{ANTISENSE['genes']}_s{ANTISENSE['seed']}")
    tx = SynthTxome(ANTISENSE["seed"], ANTISENSE["genes"])
    fa2 = os.path.join(d, "t_antisense.fasta")
    if not os.path.exists(os.path.join(d, "idx", "header.json")):
        os.makedirs(d, exist_ok=True)
        fa = os.path.join(d, "t.fasta")
        tx.write_fasta(fa)
        names, seqs = [], []
        with open(fa) as f:
            for line in f:
                if line.startswith(">"):
                    names.append(line[1:].strip())
                    seqs.append("")
                else:
                    seqs[-1] += line.strip()
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        rng = random.Random(5)
        with open(fa2, "w") as g:
            for i, (nm, s) in enumerate(zip(names, seqs)):
                g.write(f">{nm}\n{s}\n")
                if i % 2 == 0 and len(s) > 500:
                    seg = list(s[60:460])
                    for p in range(37, len(seg), rng.choice([41, 53, 67])):
                        seg[p] = rng.choice([c for c in "ACGT" if c != seg[p]])
                    rcseg = "".join(comp[c] for c in reversed(seg))
                    flank = "".join(rng.choice("ACGT") for _ in range(120))
                    g.write(f">{nm}_as\n{flank}{rcseg}{flank[::-1]}\n")
        build_index(fa2, os.path.join(d, "idx"))
    return os.path.join(d, "idx") + "/", tx, fa2


def make_bigsa_copy(src: str, dst: str) -> str:
    """Rewrites a 32-bit index directory as the BigSA flavour the reference writes for texts beyond 2^31 (IndexT = int64_t,
    src/RapMapSAIndexer.cpp:86-220): 8-byte suffix-array entries, transcript offsets, hash intervals ({u64, i64, i64} records)
    and FrugalBooMap starts / overflow pairs; header.json says BigSA.  The reference takes its int64 code path on the result
    (src/RapMapSAMapper.cpp:1210-1240); tests/test_oracle_golden.py checks that it prints the same SAM as on the original."""
    import json
    import shutil
    import struct

    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(src, "header.json")) as f:
        hdr = json.load(f)
    hdr["value0"]["BigSA"] = True
    with open(os.path.join(dst, "header.json"), "w") as f:
        json.dump(hdr, f, indent=4)
    shutil.copy(os.path.join(src, "rsd.bin"), os.path.join(dst, "rsd.bin"))
    b = open(os.path.join(src, "sa.bin"), "rb").read()
    n = struct.unpack_from("<Q", b, 0)[0]
    with open(os.path.join(dst, "sa.bin"), "wb") as f:
        f.write(struct.pack("<Q", n))
        np.frombuffer(b, dtype="<i4", offset=8, count=n).astype("<i8").tofile(f)
    b = open(os.path.join(src, "txpInfo.bin"), "rb").read()
    nt = struct.unpack_from("<Q", b, 0)[0]
    p = 8
    for _ in range(nt):
        p += 8 + struct.unpack_from("<Q", b, p)[0]
    no = struct.unpack_from("<Q", b, p)[0]
    offs = np.frombuffer(b, dtype="<i4", offset=p + 8, count=no)
    with open(os.path.join(dst, "txpInfo.bin"), "wb") as f:
        f.write(b[: p + 8])
        offs.astype("<i8").tofile(f)
        f.write(b[p + 8 + 4 * no:])

    def widen_table(blob: bytes, at: int, key_bytes: int) -> bytes:
        """sparsepp table at blob[at:]: be32 magic / table size / #records (no 64-bit escape at these sizes), group bitmaps, records
        {key, i32, i32} -> {key, i64, i64}."""
        magic, table, cnt = struct.unpack_from(">III", blob, at)
        assert magic == 0x24687531 and table != 0xFFFFFFFF and cnt != 0xFFFFFFFF
        rec = key_bytes + 8
        recs_at = len(blob) - cnt * rec
        recs = np.frombuffer(blob, dtype=np.dtype([("k", f"V{key_bytes}"), ("a", "<i4"), ("b", "<i4")]), offset=recs_at, count=cnt)
        wide = np.zeros(cnt, dtype=np.dtype([("k", f"V{key_bytes}"), ("a", "<i8"), ("b", "<i8")]))
        wide["k"], wide["a"], wide["b"] = recs["k"], recs["a"], recs["b"]
        return blob[at:recs_at] + wide.tobytes()

    if os.path.exists(os.path.join(src, "hash.bin")):
        blob = open(os.path.join(src, "hash.bin"), "rb").read()
        with open(os.path.join(dst, "hash.bin"), "wb") as f:
            f.write(widen_table(blob, 0, 8))
    else:
        shutil.copy(os.path.join(src, "hash_info.bph"), os.path.join(dst, "hash_info.bph"))
        blob = open(os.path.join(src, "hash_info.val"), "rb").read()
        nd = struct.unpack_from("<Q", blob, 0)[0]
        data = np.frombuffer(blob, dtype="<i4", offset=8, count=nd)
        lens_at = 8 + 4 * nd
        nl = struct.unpack_from("<Q", blob, lens_at)[0]
        table_at = lens_at + 8 + nl
        # overflow_: keys are IndexT too -> {i32, i32} records become {i64, i64}
        magic, table, cnt = struct.unpack_from(">III", blob, table_at)
        recs_at = len(blob) - cnt * 8
        recs = np.frombuffer(blob, dtype="<i4", offset=recs_at, count=2 * cnt)
        with open(os.path.join(dst, "hash_info.val"), "wb") as f:
            f.write(struct.pack("<Q", nd))
            data.astype("<i8").tofile(f)
            f.write(blob[lens_at:table_at])
            f.write(blob[table_at:recs_at])
            recs.astype("<i8").tofile(f)
    return dst if dst.endswith("/") else dst + "/"


# ----------------------------------------------------------------------------------------------
# golden fixtures
# ----------------------------------------------------------------------------------------------
def read_fastq(path: str):
    op = gzip.open if path.endswith(".gz") else open
    names, seqs = [], []
    with op(path, "rt") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip("\n")
            f.readline()
            f.readline()
            names.append(h[1:].rstrip("\n").split()[0])
            seqs.append(s)
    return names, seqs


def pack_fixed(seqs, L=None):
    L = L or len(seqs[0])
    a = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(len(seqs), L).copy()
    return a


def pack_ragged(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy() if off[-1] > 0 else np.zeros(1, dtype=np.uint8)
    return buf, off


def golden_sample():
    """(index dir, mate1 array, mate2 array, read length) of the committed sample_data fixture."""
    _, a = read_fastq(os.path.join(GOLD, "sample_reads_1.fastq.gz"))
    _, b = read_fastq(os.path.join(GOLD, "sample_reads_2.fastq.gz"))
    return os.path.join(GOLD, "sample_idx") + "/", pack_fixed(a), pack_fixed(b), len(a[0])


def md5(data: bytes) -> str:
    return hashlib.md5(data).hexdigest()


def explain_mismatch(res: rb.BatchResult, ref: rb.BatchResult, limit: int = 5) -> str:
    """Human-readable first differences between two batch results."""
    out = []
    n = len(ref.pair_offsets) - 1
    for i in range(n):
        a = res.hits[int(res.pair_offsets[i]) : int(res.pair_offsets[i + 1])]
        b = ref.hits[int(ref.pair_offsets[i]) : int(ref.pair_offsets[i + 1])]
        if len(a) != len(b) or not np.array_equal(a, b):
            out.append(f"pair {i}: got {a.tolist()} expected {b.tolist()}")
            if len(out) >= limit:
                break
    return "\n".join(out)
