#!/usr/bin/env python
"""Chunk-size sweep of the C-ABI call (the reference's own chunk is 10,000 pairs, src/RapMapSAMapper.cpp:853).

For each chunk size: pairs/s with device-resident buffers (one call at a time) and end to end with pinned host buffers through
rapmap_cuda_map_batch_async / _wait (one host thread, one mapper, three chunks in flight), default flags and -s, on the
benchmark index.  Output: one JSON document on stdout (kept under profiles/)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

import bench  # noqa: E402  (index cache + constants)
import rapmap_b200 as rb  # noqa: E402
from helpers import SynthTxome  # noqa: E402


def main():
    genes = int(os.environ.get("GENES", "37000"))
    sizes = [int(x) for x in os.environ.get("SIZES", "10000,65536,262144,1048576").split(",")]
    idx_dir = bench.ensure_index(genes, True)
    index = rb.Index(idx_dir, 0)
    tx = SynthTxome(bench.TX_SEED, genes)
    nmax = max(sizes)
    h1 = torch.empty((nmax, 100), dtype=torch.uint8).pin_memory()
    h2 = torch.empty((nmax, 100), dtype=torch.uint8).pin_memory()
    tx.reads(nmax, rseed=bench.READ_SEED, first=0, read_len=100, out1=h1.numpy(), out2=h2.numpy())
    d1, d2 = h1.cuda(), h2.cuda()
    out = {"index": f"{index.num_transcripts} transcripts", "rows": []}
    for sel in (False, True):
        opts = rb.default_opts(sel_aln=sel)
        for n in sizes:
            mapper = rb.Mapper(index, opts, max_batch=n, max_read_len=100)
            cap = 8 * n
            dh, do = torch.empty(cap * 28, dtype=torch.uint8, device="cuda"), torch.empty(n + 1, dtype=torch.int64, device="cuda")
            hh = [(torch.empty(cap * 28, dtype=torch.uint8).pin_memory(), torch.empty(n + 1, dtype=torch.int64).pin_memory()) for _ in range(3)]
            calls = max(8, min(400, (4 << 20) // n))
            chunks = max(1, nmax // n)

            def res(c):
                o = (c % chunks) * n
                mapper.map_batch(d1[o:o + n], d2[o:o + n], n=n, fixed_len=100, location=rb.LOC_DEVICE, hits_out=dh, offsets_out=do, out_location=rb.LOC_DEVICE, capacity=cap)

            for c in range(3):
                res(c)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for c in range(calls):
                res(c)
            torch.cuda.synchronize()
            t_res = time.perf_counter() - t0

            def e2e(count):
                for c in range(count):
                    if mapper.in_flight == 3:
                        mapper.wait()
                    o = (c % chunks) * n
                    mapper.map_batch_async(h1[o:o + n], h2[o:o + n], n=n, fixed_len=100, hits_out=hh[c % 3][0], offsets_out=hh[c % 3][1], capacity=cap)
                while mapper.in_flight:
                    mapper.wait()

            e2e(4)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e(calls)
            torch.cuda.synchronize()
            t_e2e = time.perf_counter() - t0
            t = mapper.timing()
            row = {"flags": "-s" if sel else "default", "pairs_per_call": n, "calls": calls, "resident_Mpairs_s": n * calls / t_res / 1e6,
                   "e2e_Mpairs_s": n * calls / t_e2e / 1e6, "resident_us_per_call": 1e6 * t_res / calls, "e2e_us_per_call": 1e6 * t_e2e / calls,
                   "launches_per_call": t.launches}
            out["rows"].append(row)
            print(row, file=sys.stderr, flush=True)
            mapper.close()
    os.write(bench._REAL_STDOUT, (json.dumps(out, indent=1) + "\n").encode())  # importing bench points fd 1 at stderr


if __name__ == "__main__":
    main()
