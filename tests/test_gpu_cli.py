"""GPU: the `rapmap_b200 quasimap` front end (tools/quasimap_main.cpp) writes the reference's SAM byte for byte."""
import gzip
import json
import os
import subprocess

import pytest

from helpers import GOLD, ROOT, md5

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "build", "bin", "rapmap_b200")
with open(os.path.join(GOLD, "golden.json")) as f:
    GOLDEN = json.load(f)


@pytest.mark.parametrize("flags,key", [([], "sample/default"), (["-s"], "sample/selaln")])
def test_cli_sample_sam_md5(tmp_path, flags, key):
    if not os.path.exists(CLI):
        pytest.skip("CLI not built")
    fq = []
    for m in (1, 2):
        p = tmp_path / f"r{m}.fastq"
        with gzip.open(os.path.join(GOLD, f"sample_reads_{m}.fastq.gz"), "rt") as f:
            p.write_text(f.read())
        fq.append(str(p))
    out = tmp_path / "o.sam"
    # small batch so that several map_batch calls are chained
    subprocess.run([CLI, "quasimap", "-i", os.path.join(GOLD, "sample_idx"), "-1", fq[0], "-2", os.path.join(GOLD, "sample_reads_2.fastq.gz"), "-o", str(out),
                    "--batch", "700", "-q"] + flags, check=True)
    assert md5(out.read_bytes()) == GOLDEN[key]["md5"]


def _synth_fastq(tmp_path):
    from helpers import SYNTH_BIN

    meta = GOLDEN["_meta"]["synth"]
    subprocess.run([SYNTH_BIN, "reads", "--genes", str(meta["genes"]), "--seed", str(meta["seed"]), "--pairs", str(meta["pairs"]), "--rseed", str(meta["rseed"]),
                    "--sub", str(meta["sub"]), "--ins", str(meta["ins"]), "--del", str(meta["del"]), "--n", str(meta["n"]),
                    "--out1", str(tmp_path / "y1.fastq"), "--out2", str(tmp_path / "y2.fastq")], check=True)
    return str(tmp_path / "y1.fastq"), str(tmp_path / "y2.fastq")


@pytest.mark.parametrize("key", sorted(k for k in GOLDEN if k.startswith("synth/") or k.startswith("synth_r/")))
def test_cli_every_flag_set_paired_and_unmated(tmp_path, key):
    """The pipelined front end (parser threads, two chunks in flight, -t formatter threads) over several small chunks: the
    reference's SAM byte for byte for every golden flag set, paired (-1/-2) and unmated (-r)."""
    if not os.path.exists(CLI):
        pytest.skip("CLI not built")
    y1, y2 = _synth_fastq(tmp_path)
    out = tmp_path / "o.sam"
    reads = ["-1", y1, "-2", y2] if key.startswith("synth/") else ["-r", y1]
    subprocess.run([CLI, "quasimap", "-i", os.path.join(GOLD, "synth_idx")] + reads + ["-o", str(out), "--batch", "400", "-t", "3", "-q"] + GOLDEN[key]["flags"], check=True)
    assert md5(out.read_bytes()) == GOLDEN[key]["md5"], key


def test_cli_tiny_hit_buffers_and_exact_multiple(tmp_path, monkeypatch):
    """Chunk size dividing the read count exactly (an empty last chunk) with -t 1, and -n (no output) reporting the counters."""
    if not os.path.exists(CLI):
        pytest.skip("CLI not built")
    y1, y2 = _synth_fastq(tmp_path)
    out = tmp_path / "o.sam"
    subprocess.run([CLI, "quasimap", "-i", os.path.join(GOLD, "synth_idx"), "-1", y1, "-2", y2, "-o", str(out), "--batch", "500", "-q", "-s"], check=True)
    assert md5(out.read_bytes()) == GOLDEN["synth/selaln"]["md5"]
    p = subprocess.run([CLI, "quasimap", "-i", os.path.join(GOLD, "synth_idx"), "-1", y1, "-2", y2, "-n", "--batch", "500"], check=True, capture_output=True, text=True)
    assert "In total saw 1500 reads." in p.stderr
