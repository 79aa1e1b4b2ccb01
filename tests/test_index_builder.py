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


def _repeat_fasta(path):
    """400 short transcripts sharing one 90-base element: its k-mers have SA intervals of 400 suffixes (>= 255: the overflow_
    table of FrugalBooMap)."""
    import random

    rng = random.Random(11)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    el = rnd(90)
    with open(path, "w") as f:
        for t in range(400):
            f.write(f">t{t}\n{rnd(rng.randrange(60, 200))}{el}{rnd(rng.randrange(60, 200))}C\n")


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("case", ["synth", "repeat_element"])
def test_perfect_hash_writer_matches_quasiindex_p(tmp_path, case):
    """tools/build_index.py, `-p` flavour: hash_info.bph (BooPHF levels + ranks) and hash_info.val (FrugalBooMap data_ / lens_ /
    overflow_) against the reference's own `quasiindex -p` on the same transcriptome, and the unmodified reference maps with
    the written files exactly as with its own.  The repeat-element case has intervals of >= 255 suffixes (overflow_ table)."""
    from build_index import build_fasta_index, build_synth_index

    oracle_lib()
    mine, mine_p = tmp_path / "mine", tmp_path / "mine_p"
    fa = tmp_path / "t.fasta"
    genes, seed = 40, 99
    if case == "synth":
        build_synth_index(str(mine), seed, genes, 0, device="cpu", verbose=False, perfect_dir=str(mine_p))
        subprocess.run([SYNTH_BIN, "txome", "--genes", str(genes), "--seed", str(seed), "--out", str(fa)], check=True)
    else:
        _repeat_fasta(fa)
        build_fasta_index(str(fa), str(mine), device="cpu", perfect_dir=str(mine_p))
    ref_p = tmp_path / "ref_p"
    subprocess.run([REF_BIN, "quasiindex", "-t", str(fa), "-i", str(ref_p), "-p", "-x", "4"], check=True, capture_output=True)
    for f in ("sa.bin", "txpInfo.bin", "rsd.bin"):
        assert (mine / f).read_bytes() == (ref_p / f).read_bytes(), f
    a, b = (mine_p / "hash_info.bph").read_bytes(), (ref_p / "hash_info.bph").read_bytes()
    assert a == b, "BooPHF level bitsets / rank samples differ from the reference's"
    va, vb = (mine_p / "hash_info.val").read_bytes(), (ref_p / "hash_info.val").read_bytes()
    n = int.from_bytes(va[:8], "little")
    assert va[: 16 + 5 * n] == vb[: 16 + 5 * n], "data_ / lens_ differ"
    # overflow_: same (start -> length) set; sparsepp's table size may differ from the writer's
    def overflow(v):
        tail = v[16 + 5 * n:]
        cnt = int.from_bytes(tail[8:12], "big")
        return cnt, np.sort(np.frombuffer(tail[len(tail) - 8 * cnt:], dtype=np.dtype([("s", "<i4"), ("l", "<i4")])), order="s")
    (ca, oa), (cb, ob) = overflow(va), overflow(vb)
    assert ca == cb and np.array_equal(oa, ob)
    if case == "repeat_element":
        assert ca > 0, "the shared element should produce intervals of >= 255 suffixes"
        r1, r2 = tmp_path / "r1.fq", tmp_path / "r2.fq"
        seqs = [l.strip() for l in open(fa) if not l.startswith(">")]
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        with open(r1, "w") as f1, open(r2, "w") as f2:
            for i, sq in enumerate(seqs[:300]):
                a0 = (i * 7) % max(1, len(sq) - 160)
                frag = sq[a0 : a0 + 160]
                m1, m2 = frag[:75], "".join(comp[c] for c in reversed(frag[-75:]))
                f1.write(f"@q{i}/1\n{m1}\n+\n{'I' * len(m1)}\n")
                f2.write(f"@q{i}/2\n{m2}\n+\n{'I' * len(m2)}\n")
    else:
        subprocess.run([SYNTH_BIN, "reads", "--genes", str(genes), "--seed", str(seed), "--pairs", "3000", "--rseed", "5",
                        "--out1", str(tmp_path / "r1.fq"), "--out2", str(tmp_path / "r2.fq")], check=True)
    outs = []
    for idx in (mine_p, ref_p, mine):
        o = tmp_path / (idx.name + ".sam")
        subprocess.run([REF_BIN, "quasimap", "-i", str(idx), "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq"), "-t", "1", "-s", "-m", "500", "-o", str(o)],
                       check=True, capture_output=True)
        outs.append(o.read_bytes())
    assert outs[0] == outs[1] == outs[2]
    assert outs[0].count(b"\n") > 1000
