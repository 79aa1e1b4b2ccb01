"""Host checks of the two-jobs-per-thread ksw2 DP (rapmap_b200/csrc/ksw_pair.cuh is host-callable; on the device the same
source runs with the native 16x2 instructions):

 * against the REFERENCE's own SSE function, compiled from where it lies under /root/reference (skipped without it):
   ~120,000 random / adversarial job pairs (noisy copies with indels, unrelated sequences, N / IUPAC, homopolymer runs,
   alignments off the main diagonal, short queries, windows cut by the transcript end, bandwidths 1..15, other scores);
 * against the oracle's restatement of ksw_extz2_sse41 (which is pinned to the reference by the -s golden SAM), always.
"""
import ctypes as C
import os
import random
import subprocess

import pytest

import rapmap_b200 as rb
from helpers import OracleMapper, golden_sample

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "ksw_pair_host.cpp")
REF = os.environ.get("RAPMAP_REFERENCE_DIR", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "ksw2pp", "ksw2_extz2_sse.c")), reason="needs the reference sources")
def test_pair_dp_equals_reference_sse_kernel(tmp_path):
    obj = tmp_path / "ksw2_ref.o"
    subprocess.run(["gcc", "-O2", "-msse4.1", "-I" + os.path.join(REF, "include"), "-c", os.path.join(REF, "src", "ksw2pp", "ksw2_extz2_sse.c"), "-o", str(obj)], check=True)
    exe = tmp_path / "ksw_pair_vs_ref"
    subprocess.run(["g++", "-O2", "-std=c++14", "-DWITH_REF", "-I" + os.path.join(REF, "include"), "-o", str(exe), SRC, str(obj)], check=True)
    out = subprocess.run([str(exe), "60000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
    scored = int(out.stdout.split("pairs scored ")[1].split(",")[0])
    assert scored > 40000, out.stdout  # the pair kernel takes the bulk, the rest is outside its geometry


def test_pair_dp_equals_oracle_restatement(tmp_path):
    so = tmp_path / "libksw_pair_host.so"
    subprocess.run(["g++", "-O2", "-std=c++14", "-shared", "-fPIC", "-o", str(so), SRC], check=True)
    L = C.CDLL(str(so))
    L.ksw_pair_host_score.restype = C.c_int
    L.ksw_pair_host_score.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int] + [C.c_int] * 6 + [C.POINTER(C.c_int32)]
    idx_dir = golden_sample()[0]
    rng = random.Random(7)
    checked = 0
    for go, ge, ma, mm, w in [(5, 3, 2, -4, 15), (4, 2, 1, -3, 7), (6, 1, 3, -6, 15)]:
        opts = rb.default_opts(sel_aln=True)
        opts.gap_open_penalty, opts.gap_extend_penalty, opts.match_score, opts.mismatch_penalty, opts.dp_bandwidth = go, ge, ma, mm, w
        om = OracleMapper(idx_dir, opts)
        for _ in range(400):
            qlen = rng.randint(30, 150)
            tlen = max(64, qlen + rng.choice([20, 20, 20, 5, -8, 12]))
            seqs = []
            for _j in range(2):
                t = [rng.choice("ACGT") for _ in range(tlen)]
                q, p = [], 0
                while len(q) < qlen:
                    u = rng.random()
                    if p >= tlen or u < 0.02:
                        q.append(rng.choice("ACGT")); p += 1
                    elif u < 0.03:
                        q.append(rng.choice("ACGTN"))
                    elif u < 0.04:
                        p += 1
                    else:
                        q.append(t[p]); p += 1
                seqs.append(("".join(q).encode(), "".join(t).encode()))
            out = (C.c_int32 * 2)()
            rc = L.ksw_pair_host_score(seqs[0][0], seqs[1][0], qlen, seqs[0][1], seqs[1][1], tlen, ma, mm, go, ge, w, 384, out)
            if rc != 1:
                continue
            for j in range(2):
                assert out[j] == om.ksw(seqs[j][0], seqs[j][1]), (go, ge, ma, mm, w, qlen, tlen, seqs[j])
            checked += 1
    assert checked > 900
