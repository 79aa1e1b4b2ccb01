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
    out["synth_r"] = (str(d / "y1.fastq"), None)  # unmated reads (-r): the mate-1 file alone
    return out


IDX = {"sample": "sample_idx", "synth": "synth_idx", "synth_p": "synth_idx_p", "synth_r": "synth_idx"}


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case, fastqs, tmp_path):
    dname, fname = case.split("/")
    r1, r2 = fastqs[dname]
    out = tmp_path / "o.sam"
    reads = ["-1", r1, "-2", r2] if r2 else ["-r", r1]
    subprocess.run([ORACLE_CLI, "-i", os.path.join(GOLD, IDX[dname])] + reads + ["-o", str(out)] + GOLDEN[case]["flags"], check=True, capture_output=True)
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


def test_sample_goldens_are_the_surveys_known_answers():
    """config 1 (all 10,000 sample pairs, -t 1): the md5s SURVEY.md section 4 records for the unmodified reference."""
    assert GOLDEN["sample/default"]["md5"] == "5271acf4e1c0e5d22b43ab7859f825d3" and GOLDEN["sample/default"]["lines"] == 28523
    assert GOLDEN["sample/selaln"]["md5"] == "ddd30824c80623c4e272324cd6e7784f" and GOLDEN["sample/selaln"]["lines"] == 28523


ANTISENSE_FLAGS = [[], ["--noSensitive"], ["--noStrictCheck"], ["--noSensitive", "--noStrictCheck"], ["-s"], ["-s", "--noSensitive"], ["-s", "--noStrictCheck"],
                   ["-s", "--noSensitive", "--noStrictCheck"], ["-z", "0.7"], ["-m", "2"], ["-f"], ["-f", "--noStrictCheck"]]


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_live_reference_on_antisense_transcriptome(tmp_path):
    """Transcriptome with diverged antisense copies (helpers.antisense_index): reads hit both strands, so --noSensitive
    (NIP skip + k-mer votes), --noStrictCheck (no strand filter) and their combinations really change the output.  The
    oracle must equal the reference under each, and each must differ from the default output (the fixture discriminates)."""
    from helpers import ANTISENSE, antisense_index

    idx, _, _ = antisense_index()
    d = tmp_path
    subprocess.run([SYNTH_BIN, "reads", "--genes", str(ANTISENSE["genes"]), "--seed", str(ANTISENSE["seed"]), "--pairs", "4000", "--rseed", "77", "--sub", "15000",
                    "--ins", "1000", "--del", "1000", "--n", "1000", "--out1", str(d / "a1.fastq"), "--out2", str(d / "a2.fastq")], check=True)
    seen = {}
    for flags in ANTISENSE_FLAGS:
        subprocess.run([REF_BIN, "quasimap", "-i", idx, "-1", str(d / "a1.fastq"), "-2", str(d / "a2.fastq"), "-t", "1", "-o", str(d / "ref.sam")] + flags,
                       check=True, capture_output=True)
        subprocess.run([ORACLE_CLI, "-i", idx, "-1", str(d / "a1.fastq"), "-2", str(d / "a2.fastq"), "-o", str(d / "ora.sam")] + flags, check=True, capture_output=True)
        ref = (d / "ref.sam").read_bytes()
        assert ref == (d / "ora.sam").read_bytes(), flags
        seen[" ".join(flags)] = md5(ref)
    assert len(set(seen.values())) == len(seen), f"flag sets with identical output: {seen}"
    # unmated reads (-r) on the same index
    for flags in ([], ["-s"], ["--noSensitive"], ["-s", "--noStrictCheck"], ["-m", "1"]):
        subprocess.run([REF_BIN, "quasimap", "-i", idx, "-r", str(d / "a1.fastq"), "-t", "1", "-o", str(d / "ref.sam")] + flags, check=True, capture_output=True)
        subprocess.run([ORACLE_CLI, "-i", idx, "-r", str(d / "a1.fastq"), "-o", str(d / "ora.sam")] + flags, check=True, capture_output=True)
        assert (d / "ref.sam").read_bytes() == (d / "ora.sam").read_bytes(), ["-r"] + flags


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("idx,case", [("synth_idx", "synth/default"), ("synth_idx", "synth/selaln"), ("synth_idx_p", "synth_p/default")])
def test_bigsa_rewrite_is_accepted_by_the_reference(fastqs, tmp_path, idx, case):
    """helpers.make_bigsa_copy turns a 32-bit index into the BigSA (int64) flavour: the unmodified reference dispatches on
    header.json (src/RapMapSAMapper.cpp:1210-1240), maps with its int64 instantiation and prints the golden SAM."""
    from helpers import make_bigsa_copy

    big = make_bigsa_copy(os.path.join(GOLD, idx), str(tmp_path / "big"))
    r1, r2 = fastqs["synth"]
    out = tmp_path / "o.sam"
    subprocess.run([REF_BIN, "quasimap", "-i", big, "-1", r1, "-2", r2, "-t", "1", "-o", str(out)] + GOLDEN[case]["flags"], check=True, capture_output=True)
    assert md5(out.read_bytes()) == GOLDEN[case]["md5"]
