"""CPU: tools/build_index.py writes the reference's index files (checked against `rapmap_ref quasiindex` and by having
the unmodified reference map with the result)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import REF_BIN, ROOT, SYNTH_BIN, have_ref, oracle_lib

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_builder_matches_quasiindex(tmp_path):
    from build_index import build_synth_index

    oracle_lib()
    genes, seed = 40, 99
    mine = tmp_path / "mine"
    build_synth_index(str(mine), seed, genes, device="cpu", verbose=False)
    fa = tmp_path / "t.fasta"
    subprocess.run([SYNTH_BIN, "txome", "--genes", str(genes), "--seed", str(seed), "--out", str(fa)], check=True)
    ref = tmp_path / "ref"
    subprocess.run([REF_BIN, "quasiindex", "-t", str(fa), "-i", str(ref)], check=True, capture_output=True)
    for f in ("sa.bin", "txpInfo.bin", "rsd.bin"):
        assert (mine / f).read_bytes() == (ref / f).read_bytes(), f
    a, b = (mine / "hash.bin").read_bytes(), (ref / "hash.bin").read_bytes()
    assert len(a) == len(b) and a[:12] == b[:12]
    nb = int.from_bytes(a[8:12], "big")
    rec = np.dtype([("k", "<u8"), ("b", "<i4"), ("e", "<i4")])
    ra = np.sort(np.frombuffer(a[-nb * 16:], dtype=rec), order="k")
    rb_ = np.sort(np.frombuffer(b[-nb * 16:], dtype=rec), order="k")
    assert np.array_equal(ra, rb_)
    # the unmodified reference maps with the builder's index exactly as with its own
    subprocess.run([SYNTH_BIN, "reads", "--genes", str(genes), "--seed", str(seed), "--pairs", "3000", "--rseed", "5", "--out1", str(tmp_path / "r1.fq"),
                    "--out2", str(tmp_path / "r2.fq")], check=True)
    outs = []
    for idx in (mine, ref):
        o = tmp_path / (idx.name + ".sam")
        subprocess.run([REF_BIN, "quasimap", "-i", str(idx), "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq"), "-t", "1", "-s", "-o", str(o)],
                       check=True, capture_output=True)
        outs.append(o.read_bytes())
    assert outs[0] == outs[1]
