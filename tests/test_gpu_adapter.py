"""GPU: the source-level adapter (include/rapmap_b200/adapter.hpp) inside the reference's own header tree.

oracle/_ref/adapter_sam is tests/cpp/adapter_sam.cpp - the stub of INTEGRATION.md section 1 - compiled against the unmodified
reference headers and linked with the reference's objects (oracle/build_adapter_harness.sh): the reference's FASTQ parser
feeds ReadGroup chunks to rapmap_b200::BatchMapper, and the std::vector<QuasiAlignment> it returns are printed by the
reference's own writeSAMHeader / writeAlignmentsToStream.  The SAM must be the golden SAM of `rapmap quasimap`."""
import gzip
import json
import os
import subprocess

import pytest

from helpers import GOLD, ROOT, SYNTH_BIN, md5

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "oracle", "_ref", "adapter_sam")

with open(os.path.join(GOLD, "golden.json")) as f:
    GOLDEN = json.load(f)
META = GOLDEN["_meta"]["synth"]


@pytest.fixture(scope="module")
def fastqs(tmp_path_factory):
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/adapter_sam not built (needs the reference sources at build time)")
    d = tmp_path_factory.mktemp("fq")
    for m in (1, 2):
        with gzip.open(os.path.join(GOLD, f"sample_reads_{m}.fastq.gz"), "rt") as f:
            (d / f"s{m}.fastq").write_text(f.read())
    subprocess.run([SYNTH_BIN, "reads", "--genes", str(META["genes"]), "--seed", str(META["seed"]), "--pairs", str(META["pairs"]), "--rseed", str(META["rseed"]),
                    "--sub", str(META["sub"]), "--ins", str(META["ins"]), "--del", str(META["del"]), "--n", str(META["n"]),
                    "--out1", str(d / "y1.fastq"), "--out2", str(d / "y2.fastq")], check=True)
    return d


@pytest.mark.parametrize("case,idx,reads,flags", [
    ("sample/default", "sample_idx", ("s1.fastq", "s2.fastq"), []),
    ("sample/selaln", "sample_idx", ("s1.fastq", "s2.fastq"), ["-s"]),
    ("synth/default", "synth_idx", ("y1.fastq", "y2.fastq"), []),
    ("synth/selaln", "synth_idx", ("y1.fastq", "y2.fastq"), ["-s"]),
    ("synth_r/default", "synth_idx", ("y1.fastq", None), []),
    ("synth_r/selaln", "synth_idx", ("y1.fastq", None), ["-s"]),
])
def test_adapter_stub_prints_the_golden_sam_through_the_reference_writers(fastqs, tmp_path, case, idx, reads, flags):
    out = tmp_path / "a.sam"
    rd = ["-1", str(fastqs / reads[0]), "-2", str(fastqs / reads[1])] if reads[1] else ["-r", str(fastqs / reads[0])]
    # one chunk holds the whole file: the order of the records is then the file order (the reference's parser threads hand
    # chunks to the consumer through a concurrent queue and do not promise their order)
    p = subprocess.run([HARNESS, os.path.join(GOLD, idx) + "/", str(out)] + flags + rd + ["--chunk", "10000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert md5(out.read_bytes()) == GOLDEN[case]["md5"], f"{case}: SAM through the reference's writers differs from the golden SAM\n{p.stderr[-500:]}"


@pytest.mark.parametrize("case,flags", [("synth/default", []), ("synth/selaln", ["-s"])])
def test_adapter_stub_with_several_ragged_chunks(fastqs, tmp_path, case, flags):
    """A chunk size that does not divide the read count (two full chunks and a ragged last one): the same records as the
    golden SAM; compared as a multiset of lines because the reference's parser may deliver the chunks in any order."""
    out = tmp_path / "a.sam"
    p = subprocess.run([HARNESS, os.path.join(GOLD, "synth_idx") + "/", str(out)] + flags + ["-1", str(fastqs / "y1.fastq"), "-2", str(fastqs / "y2.fastq"),
                        "--chunk", "700"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    with gzip.open(os.path.join(GOLD, case.replace("/", "_") + ".sam.gz"), "rb") as f:
        gold = f.read()
    assert md5(gold) == GOLDEN[case]["md5"]
    assert sorted(out.read_bytes().split(b"\n")) == sorted(gold.split(b"\n"))
