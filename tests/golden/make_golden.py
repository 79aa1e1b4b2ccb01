#!/usr/bin/env python
"""Regenerates tests/golden/ from the compiled reference (oracle/_ref/rapmap_ref, built by oracle/build_ref.sh).

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py
  sample_idx/            quasiindex of the reference's sample_data/transcripts.fasta (15 transcripts)
  sample_reads_{1,2}.fastq.gz   all 10,000 pairs of sample_data/reads_{1,2}.fastq (2 x 50 bp, error free): the md5s of
                         sample/default and sample/selaln are the known answers of SURVEY.md §4
  synth_idx/, synth_idx_p/      quasiindex (dense, -p) of `synth txome --genes 8 --seed 777`
  golden.json            md5 of `rapmap_ref quasimap -t 1 <flags>` SAM for every (dataset, flag set)
  *.sam.gz               full SAM for the default and -s flag sets (for readable diffs)
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "rapmap_ref")
SYNTH = os.path.join(ROOT, "build", "bin", "synth")
REFDATA = os.environ.get("RAPMAP_REFERENCE_DIR", "/root/reference") + "/sample_data"

FLAGSETS = {
    "default": [],
    "selaln": ["-s"],
    "selaln_hard": ["-s", "--hardFilter"],
    "selaln_bt2": ["--mimicBT2"],
    "selaln_strictbt2": ["--mimicStrictBT2"],
    "selaln_w5": ["-s", "--dpBandwidth", "5"],
    "selaln_scores": ["-s", "--ma", "3", "--mm", "-2", "--go", "5", "--ge", "1", "--minScoreFrac", "0.5"],
    "selaln_noorph_nodove": ["-s", "--noOrphans", "--noDovetail"],
    "selaln_slack0": ["-s", "--consensusSlack", "0"],
    "selaln_ext3": ["-s", "--maxMMPExtension", "3"],
    "fuzzy": ["-f"],
    "noorphans_nodovetail": ["--noOrphans", "--noDovetail"],
    "maxhits2": ["-m", "2"],
    "cov07": ["-z", "0.7"],
    "nosensitive": ["--noSensitive"],
    "nostrict": ["--noStrictCheck"],
    "selaln_recover": ["-s", "--recoverOrphans"],
    "fuzzy_recover": ["-f", "--recoverOrphans"],
}
SYNTH_PAIRS = 1500
KEEP_SAM = {"default", "selaln"}


def run(cmd):
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def index(fasta, out, perfect=False):
    if os.path.exists(out):
        shutil.rmtree(out)
    tmp = tempfile.mkdtemp()
    run([REF, "quasiindex", "-t", fasta, "-i", tmp + "/idx"] + (["-p", "-x", "2"] if perfect else []))
    os.makedirs(out)
    keep = ["header.json", "sa.bin", "txpInfo.bin", "rsd.bin"] + (["hash_info.bph", "hash_info.val"] if perfect else ["hash.bin"])
    for f in keep:
        shutil.copy(os.path.join(tmp, "idx", f), os.path.join(out, f))
    shutil.rmtree(tmp)


def head_fastq(src, dst, pairs):
    with open(src) as f, gzip.open(dst, "wt", compresslevel=9) as g:
        for i, line in enumerate(f):
            if i >= 4 * pairs:
                break
            g.write(line)


def quasimap(idx, r1, r2, flags):
    with tempfile.NamedTemporaryFile(suffix=".sam") as t:
        reads = ["-1", r1, "-2", r2] if r2 else ["-r", r1]
        run([REF, "quasimap", "-i", idx] + reads + ["-t", "1", "-o", t.name] + flags)
        return open(t.name, "rb").read()


# unmated reads (quasimap -r: processReadsSingleSA): mate-1 reads of the synthetic set
UNMATED_FLAGSETS = {"default": [], "selaln": ["-s"], "selaln_hard": ["-s", "--hardFilter"], "maxhits2": ["-m", "2"], "nosensitive": ["--noSensitive"]}


def main():
    assert os.path.exists(REF), "build oracle/_ref first (oracle/build_ref.sh)"
    golden = {}
    tmp = tempfile.mkdtemp()
    # ---- sample_data
    index(os.path.join(REFDATA, "transcripts.fasta"), os.path.join(HERE, "sample_idx"))
    head_fastq(os.path.join(REFDATA, "reads_1.fastq"), os.path.join(HERE, "sample_reads_1.fastq.gz"), 10000)
    head_fastq(os.path.join(REFDATA, "reads_2.fastq"), os.path.join(HERE, "sample_reads_2.fastq.gz"), 10000)
    for m in (1, 2):
        with gzip.open(os.path.join(HERE, f"sample_reads_{m}.fastq.gz"), "rt") as f, open(os.path.join(tmp, f"s{m}.fastq"), "w") as g:
            g.write(f.read())
    # ---- synthetic
    fa = os.path.join(tmp, "synth.fasta")
    run([SYNTH, "txome", "--genes", "8", "--seed", "777", "--out", fa])
    index(fa, os.path.join(HERE, "synth_idx"))
    index(fa, os.path.join(HERE, "synth_idx_p"), perfect=True)
    run([SYNTH, "reads", "--genes", "8", "--seed", "777", "--pairs", str(SYNTH_PAIRS), "--rseed", "4242", "--sub", "20000", "--ins", "2000",
         "--del", "2000", "--n", "2000", "--out1", os.path.join(tmp, "y1.fastq"), "--out2", os.path.join(tmp, "y2.fastq")])
    datasets = {
        "sample": (os.path.join(HERE, "sample_idx"), os.path.join(tmp, "s1.fastq"), os.path.join(tmp, "s2.fastq")),
        "synth": (os.path.join(HERE, "synth_idx"), os.path.join(tmp, "y1.fastq"), os.path.join(tmp, "y2.fastq")),
        "synth_p": (os.path.join(HERE, "synth_idx_p"), os.path.join(tmp, "y1.fastq"), os.path.join(tmp, "y2.fastq")),
        "synth_r": (os.path.join(HERE, "synth_idx"), os.path.join(tmp, "y1.fastq"), None),
    }
    for dname, (idx, r1, r2) in datasets.items():
        for fname, flags in (UNMATED_FLAGSETS if dname == "synth_r" else FLAGSETS).items():
            if dname not in ("synth", "synth_r") and fname not in ("default", "selaln"):
                continue
            sam = quasimap(idx, r1, r2, flags)
            golden[f"{dname}/{fname}"] = {"flags": flags, "md5": hashlib.md5(sam).hexdigest(), "lines": sam.count(b"\n")}
            if fname in KEEP_SAM and dname == "synth":
                with gzip.open(os.path.join(HERE, f"{dname}_{fname}.sam.gz"), "wb", compresslevel=9) as g:
                    g.write(sam)
            print(dname, fname, golden[f"{dname}/{fname}"]["md5"], golden[f"{dname}/{fname}"]["lines"])
    golden["_meta"] = {"synth": {"genes": 8, "seed": 777, "pairs": SYNTH_PAIRS, "rseed": 4242, "sub": 20000, "ins": 2000, "del": 2000, "n": 2000},
                       "reference": "COMBINE-lab/RapMap v0.6.0 (af025a4), g++ -O3 -std=c++14 -ffp-contract=off"}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    sys.exit(main())
