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
