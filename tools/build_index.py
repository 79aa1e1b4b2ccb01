"""Fast construction of a reference-format RapMapSAIndex directory for SYNTHETIC transcriptomes.

Why: the reference's `quasiindex` needs ~6 minutes for the 200k-transcript benchmark index (one thread
inserting 9e7 k-mers into sparsepp); bench.py cannot spend that on every run.  This tool writes the SAME
files (header.json, txpInfo.bin, rsd.bin, sa.bin, hash.bin — formats of SURVEY.md §5.2) in seconds:
suffix array by prefix doubling with torch sorts (on the GPU when there is one), k-mer intervals by one
vectorised pass over SA order, hash.bin laid out with sparsepp's own bucket placement (tools/synth.cpp)
so that the UNMODIFIED reference binary loads it too.  tests/test_index_builder.py checks the files
against `rapmap_ref quasiindex` output (sa.bin / txpInfo.bin / rsd.bin byte-identical, hash.bin identical
as a record set).  Index construction itself is out of the hot-path scope: this is
data tooling for tests and bench.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def suffix_array(text: torch.Tensor) -> torch.Tensor:
    """Plain byte-order suffix array (shorter suffix first on ties) of a uint8 tensor; int64 result."""
    dev = text.device
    n = text.numel()
    lut = torch.zeros(256, dtype=torch.int64, device=dev)
    for i, ch in enumerate(sorted(set(text.unique().tolist()))):
        lut[ch] = i + 1  # 0 = past the end (smallest)
    nsym = int(lut.max().item()) + 1
    bits = max(1, (nsym - 1).bit_length())
    h = max(1, 62 // bits)
    code = torch.zeros(n + h, dtype=torch.int64, device=dev)
    code[:n] = lut[text.long()]
    key = torch.zeros(n, dtype=torch.int64, device=dev)
    for j in range(h):
        key |= code[j : j + n] << (bits * (h - 1 - j))
    del code

    def rerank(key):
        skey, perm = torch.sort(key)
        flag = torch.ones(n, dtype=torch.int64, device=dev)
        flag[1:] = (skey[1:] != skey[:-1]).long()
        del skey
        r = torch.cumsum(flag, 0)
        del flag
        rank = torch.empty(n, dtype=torch.int64, device=dev)
        rank[perm] = r
        mx = int(r[-1].item())
        return rank, perm, mx

    rank, perm, mx = rerank(key)
    del key
    while mx < n:
        nxt = torch.zeros(n, dtype=torch.int64, device=dev)
        if h < n:
            nxt[: n - h] = rank[h:]
        key = rank * (n + 1) + nxt
        del nxt
        rank, perm, mx = rerank(key)
        del key
        h *= 2
    return perm


def kmer_intervals(text: torch.Tensor, sa: torch.Tensor, k: int):
    """(kmer word, begin, end) of every maximal run of SA entries sharing a valid k-mer (reference buildHash)."""
    dev = text.device
    n = text.numel()
    lut = torch.full((256,), -1, dtype=torch.int64, device=dev)
    for ch, c in ((65, 0), (67, 1), (71, 2), (84, 3)):
        lut[ch] = c
    c2 = torch.zeros(n + k, dtype=torch.int64, device=dev)
    c2[:n] = lut[text.long()]
    bad = (c2 < 0).long()
    bad[n:] = 1
    c2.clamp_(min=0)
    word = torch.zeros(n, dtype=torch.int64, device=dev)
    for j in range(k):
        word |= c2[j : j + n] << (2 * (k - 1 - j))
    cs = torch.zeros(n + k + 1, dtype=torch.int64, device=dev)
    cs[1:] = torch.cumsum(bad, 0)
    valid = (cs[k : k + n] - cs[:n]) == 0
    del c2, bad, cs
    w = word[sa]
    v = valid[sa]
    del word, valid
    start = v.clone()
    start[1:] &= (~v[:-1]) | (w[1:] != w[:-1])
    end = v.clone()
    end[:-1] &= (~v[1:]) | (w[1:] != w[:-1])
    b = torch.nonzero(start).squeeze(1)
    e = torch.nonzero(end).squeeze(1) + 1
    return w[b], b, e


def _header(perfect: bool, k: int) -> dict:
    return {"value0": {"IndexType": 1, "IndexVersion": "q5", "UsesKmers": True, "KmerLen": k, "BigSA": False, "PerfectHash": perfect,
                       "SeqHash": "", "NameHash": "", "SeqHash512": "", "NameHash512": ""}}


def _synth_lib():
    from helpers import synth_lib

    L = synth_lib()
    L.synth_txome_concat.restype = C.c_int64
    L.synth_txome_concat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.synth_txome_name.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
    L.synth_write_dense_hash.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p]
    L.synth_write_perfect_hash.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p]
    return L


def write_index(out_dir: str, text_np: np.ndarray, starts: np.ndarray, clens: np.ndarray, names: list, k: int = 31, device: str | None = None,
                perfect_dir: str | None = None) -> dict:
    """The index files for a concatenated transcript text (upper-case ACGT.. with '$' after every transcript, laid out as
    src/RapMapSAIndexer.cpp:536-635 does).  Dense flavour to out_dir; with perfect_dir also the `-p` flavour of the SAME index
    (hash_info.bph / hash_info.val from tools/synth.cpp's BooPHF + FrugalBooMap writer; sa.bin, txpInfo.bin and rsd.bin are
    links to out_dir's)."""
    L = _synth_lib()
    n, ntxp = int(text_np.size), len(names)
    assert n > 0 and n + 1 < 2**31, "text too long for a 32-bit index"
    dev = device or ("cuda" if torch.cuda.is_available() else "cpu")
    text = torch.from_numpy(text_np).to(dev)
    t1 = time.time()
    sa = suffix_array(text)
    if dev == "cuda":
        torch.cuda.synchronize()
    t2 = time.time()
    keys, b, e = kmer_intervals(text, sa, k)
    keys_np = keys.cpu().numpy().view(np.uint64)
    b_np = b.to(torch.int32).cpu().numpy()
    e_np = e.to(torch.int32).cpu().numpy()
    sa_np = sa.to(torch.int32).cpu().numpy()
    del sa, keys, b, e, text
    if dev == "cuda":
        torch.cuda.empty_cache()
    t3 = time.time()

    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "sa.bin"), "wb") as f:
        f.write(struct.pack("<Q", n))
        sa_np.tofile(f)
    with open(os.path.join(out_dir, "txpInfo.bin"), "wb") as f:
        f.write(struct.pack("<Q", ntxp))
        for nm in names:
            bnm = nm.encode()
            f.write(struct.pack("<Q", len(bnm)))
            f.write(bnm)
        f.write(struct.pack("<Q", ntxp))
        starts.astype(np.int32).tofile(f)
        f.write(struct.pack("<Q", n))
        text_np.tofile(f)
        f.write(struct.pack("<Q", ntxp))
        clens.astype(np.uint32).tofile(f)
    with open(os.path.join(out_dir, "rsd.bin"), "wb") as f:
        f.write(struct.pack("<Q", n))
        np.packbits(text_np == ord("$"), bitorder="little").tofile(f)
    rc = L.synth_write_dense_hash(keys_np.ctypes.data, b_np.ctypes.data, e_np.ctypes.data, len(keys_np), os.fsencode(os.path.join(out_dir, "hash.bin")))
    assert rc == 0, rc
    with open(os.path.join(out_dir, "header.json"), "w") as f:
        json.dump(_header(False, k), f, indent=4)
    t4 = time.time()
    if perfect_dir:
        os.makedirs(perfect_dir, exist_ok=True)
        rc = L.synth_write_perfect_hash(keys_np.ctypes.data, b_np.ctypes.data, e_np.ctypes.data, len(keys_np), os.fsencode(os.path.join(perfect_dir, "hash_info")))
        assert rc == 0, rc
        for name in ("sa.bin", "txpInfo.bin", "rsd.bin"):
            dst = os.path.join(perfect_dir, name)
            if os.path.lexists(dst):
                os.remove(dst)
            os.symlink(os.path.relpath(os.path.join(out_dir, name), perfect_dir), dst)
        with open(os.path.join(perfect_dir, "header.json"), "w") as f:
            json.dump(_header(True, k), f, indent=4)
    t5 = time.time()
    return {"n": n, "ntxp": ntxp, "kmers": int(len(keys_np)), "device": dev, "s_sa": t2 - t1, "s_kmers": t3 - t2, "s_write": t4 - t3, "s_perfect": t5 - t4}


def build_synth_index(out_dir: str, seed: int, genes: int, repeats: int = 0, k: int = 31, device: str | None = None, verbose: bool = True,
                      perfect_dir: str | None = None) -> dict:
    """Index of the synthetic transcriptome (tools/synth.cpp) `seed, genes, repeats`."""
    from helpers import SynthTxome

    t0 = time.time()
    L = _synth_lib()
    tx = SynthTxome(seed, genes, repeats)
    ntxp = tx.ntxp
    tlen = L.synth_txome_text_len(tx.h)
    buf = np.empty(tlen + ntxp, dtype=np.uint8)
    starts = np.empty(ntxp, dtype=np.int64)
    clens = np.empty(ntxp, dtype=np.uint32)
    n = L.synth_txome_concat(tx.h, buf.ctypes.data, starts.ctypes.data, clens.ctypes.data)
    namebuf = C.create_string_buffer(ntxp * 24 + 16)
    L.synth_txome_name(seed, genes, repeats, namebuf, len(namebuf))
    names = namebuf.value.decode().split("\n")[:-1]
    assert len(names) == ntxp
    info = write_index(out_dir, buf[:n], starts, clens, names, k, device, perfect_dir)
    info["s_total"] = time.time() - t0
    if verbose:
        print("[build_index]", json.dumps(info), file=sys.stderr)
    return info


def build_fasta_index(fasta: str, out_dir: str, k: int = 31, device: str | None = None, perfect_dir: str | None = None) -> dict:
    """Index of a plain FASTA whose records need none of the indexer's clean-ups: upper-case ACGT only, no poly-A tail of >= 10
    bases, names without spaces, every transcript longer than k (src/RapMapSAIndexer.cpp:536-635 would otherwise replace Ns,
    clip tails, cut names and drop short records - this tool refuses such input instead of approximating)."""
    names, seqs = [], []
    with open(fasta) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                names.append(line[1:])
                seqs.append([])
            elif line:
                seqs[-1].append(line)
    seqs = ["".join(x) for x in seqs]
    for nm, sq in zip(names, seqs):
        if " " in nm or "\t" in nm or set(sq) - set("ACGT") or len(sq) <= k or sq.endswith("A" * 10):
            raise ValueError(f"record {nm!r} needs the reference indexer's clean-up rules; use rapmap quasiindex")
    text = ("$".join(seqs) + "$").encode()
    starts = np.zeros(len(seqs), dtype=np.int64)
    starts[1:] = np.cumsum([len(sq) + 1 for sq in seqs])[:-1]
    clens = np.array([len(sq) for sq in seqs], dtype=np.uint32)
    return write_index(out_dir, np.frombuffer(text, dtype=np.uint8).copy(), starts, clens, names, k, device, perfect_dir)


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--genes", type=int, default=37000)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--repeats", type=int, default=0)
    ap.add_argument("--perfect-out", default=None, help="also write the -p (BooPHF) flavour of the index to this directory")
    a = ap.parse_args()
    build_synth_index(a.out, a.seed, a.genes, a.repeats, perfect_dir=a.perfect_out)
