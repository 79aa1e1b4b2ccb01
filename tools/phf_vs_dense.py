#!/usr/bin/env python
"""Dense vs -p (BooPHF) index on the same 13.9k-transcript synthetic transcriptome: SA-lookup kernel time for one batch.
A data point for BASELINE configs[3]; the 203k-transcript -p index takes the reference's quasiindex minutes to build,
so the comparison is made at the size the parity tests use.  Needs a GPU and oracle/_ref (index construction)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import rapmap_b200 as rb  # noqa: E402
from helpers import synth_index  # noqa: E402

n = 1 << 18
out = {}
d_dir, tx = synth_index(2500)
p_dir, _ = synth_index(2500, perfect=True)
s1, s2 = tx.reads(n, rseed=5)
ref = None
for name, d in (("dense", d_dir), ("perfect_hash", p_dir)):
    idx = rb.Index(d, 0)
    m = rb.Mapper(idx, rb.default_opts(), max_batch=n, max_read_len=100)
    for _ in range(3):
        res = m.map_batch(s1, s2, n=n, fixed_len=100)
    t = m.timing()
    out[name] = {"pairs": n, "ms_sa_collect": t.ms_sa_collect, "ms_total": t.ms_total, "hits": int(res.num_hits), "index_bytes": int(idx.device_bytes)}
    if ref is None:
        ref = res.hits.copy()
    else:
        out["identical_hits"] = bool(np.array_equal(ref, res.hits))
print(json.dumps(out))
