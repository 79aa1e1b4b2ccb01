"""CPU: the oracle restatement reproduces the reference's own output (golden vectors made by
tests/golden/make_golden.py with the compiled, unmodified reference) byte for byte, for every flag set."""
import gzip
import json
import os
import subprocess

import pytest

from helpers import GOLD, ORACLE_CLI, REF_BIN, SYNTH_BIN, have_ref, md5, oracle_lib

with open(os.path.join(GOLD, "golden.json")) as f:
    GOLDEN = json.load(f)
META = GOLDEN["_meta"]["synth"]
CASES = sorted(k for k in GOLDEN if not k.startswith("_"))


@pytest.fixture(scope="module")
def fastqs(tmp_path_factory):
    oracle_lib()  # builds oracle/_build if needed
    d = tmp_path_factory.mktemp("fq")
    out = {}
    for m in (1, 2):
        p = d / f"s{m}.fastq"
        with gzip.open(os.path.join(GOLD, f"sample_reads_{m}.fastq.gz"), "rt") as f:
            p.write_text(f.read())
    out["sample"] = (str(d / "s1.fastq"), str(d / "s2.fastq"))
    if not os.path.exists(SYNTH_BIN):
        os.makedirs(os.path.dirname(SYNTH_BIN), exist_ok=True)
        subprocess.run(["g++", "-O2", "-fopenmp", "-DSYNTH_MAIN", os.path.join(os.path.dirname(GOLD), "..", "tools", "synth.cpp"), "-o", SYNTH_BIN], check=True)
    subprocess.run([SYNTH_BIN, "reads", "--genes", str(META["genes"]), "--seed", str(META["seed"]), "--pairs", str(META["pairs"]), "--rseed", str(META["rseed"]),
                    "--sub", str(META["sub"]), "--ins", str(META["ins"]), "--del", str(META["del"]), "--n", str(META["n"]),
                    "--out1", str(d / "y1.fastq"), "--out2", str(d / "y2.fastq")], check=True)
    out["synth"] = out["synth_p"] = (str(d / "y1.fastq"), str(d / "y2.fastq"))
    return out


IDX = {"sample": "sample_idx", "synth": "synth_idx", "synth_p": "synth_idx_p"}


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case, fastqs, tmp_path):
    dname, fname = case.split("/")
    r1, r2 = fastqs[dname]
    out = tmp_path / "o.sam"
    subprocess.run([ORACLE_CLI, "-i", os.path.join(GOLD, IDX[dname]), "-1", r1, "-2", r2, "-o", str(out)] + GOLDEN[case]["flags"], check=True, capture_output=True)
    sam = out.read_bytes()
    if md5(sam) != GOLDEN[case]["md5"]:
        gz = os.path.join(GOLD, f"{dname}_{fname}.sam.gz")
        detail = ""
        if os.path.exists(gz):
            exp = gzip.open(gz, "rb").read().split(b"\n")
            got = sam.split(b"\n")
            for i, (a, b) in enumerate(zip(got, exp)):
                if a != b:
                    detail = f"\nfirst diff at line {i}:\n got {a[:200]!r}\n exp {b[:200]!r}"
                    break
        pytest.fail(f"{case}: oracle SAM md5 {md5(sam)} != reference {GOLDEN[case]['md5']}{detail}")
    assert sam.count(b"\n") == GOLDEN[case]["lines"]


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_live_reference_noisy_reads(fastqs, tmp_path):
    """Fresh comparison against the reference binary on a different, noisier read set (both -s and default)."""
    d = tmp_path
    subprocess.run([SYNTH_BIN, "reads", "--genes", "8", "--seed", "777", "--pairs", "2500", "--rseed", "99", "--sub", "40000", "--ins", "5000", "--del", "5000",
                    "--n", "8000", "--out1", str(d / "a1.fastq"), "--out2", str(d / "a2.fastq")], check=True)
    for flags in ([], ["-s"], ["-s", "--dpBandwidth", "3"], ["--noSensitive"], ["--noStrictCheck"], ["--noSensitive", "--noStrictCheck"],
                  ["-s", "--noSensitive"], ["-f", "--noStrictCheck"], ["-s", "--recoverOrphans"], ["-f", "--recoverOrphans"],
                  ["-s", "--recoverOrphans", "--hardFilter"]):
        subprocess.run([REF_BIN, "quasimap", "-i", os.path.join(GOLD, "synth_idx"), "-1", str(d / "a1.fastq"), "-2", str(d / "a2.fastq"), "-t", "1", "-o", str(d / "ref.sam")] + flags,
                       check=True, capture_output=True)
        subprocess.run([ORACLE_CLI, "-i", os.path.join(GOLD, "synth_idx"), "-1", str(d / "a1.fastq"), "-2", str(d / "a2.fastq"), "-o", str(d / "ora.sam")] + flags,
                       check=True, capture_output=True)
        assert (d / "ref.sam").read_bytes() == (d / "ora.sam").read_bytes(), flags
