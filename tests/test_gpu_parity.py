"""GPU: the CUDA path (through the C-ABI) is bit-identical to the CPU oracle and to the reference's golden SAM."""
import gzip
import json
import os

import numpy as np
import pytest

import rapmap_b200 as rb
from helpers import (GOLD, OracleMapper, SynthTxome, explain_mismatch, golden_sample, have_ref, md5, pack_fixed, pack_ragged, read_fastq,
                     synth_index)

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLD, "golden.json")) as f:
    GOLDEN = json.load(f)
META = GOLDEN["_meta"]["synth"]


def assert_same(res, ref, what=""):
    assert res.num_hits == ref.num_hits, f"{what}: hits {res.num_hits} != {ref.num_hits}\n" + explain_mismatch(res, ref)
    assert np.array_equal(res.pair_offsets, ref.pair_offsets), f"{what}: offsets differ\n" + explain_mismatch(res, ref)
    assert np.array_equal(res.hits, ref.hits), f"{what}: records differ\n" + explain_mismatch(res, ref)
    assert np.array_equal(res.counters, ref.counters), f"{what}: counters {res.counters} != {ref.counters}"


def make_mapper(index, opts, n, L):
    try:
        return rb.Mapper(index, opts, max_batch=n, max_read_len=L)
    except rb.RapMapCudaError as e:
        if e.code == rb.ERR_UNSUPPORTED:
            pytest.skip(f"refused by the device path: {e}")
        raise


@pytest.fixture(scope="module")
def sample():
    idx_dir, s1, s2, L = golden_sample()
    n1, _ = read_fastq(os.path.join(GOLD, "sample_reads_1.fastq.gz"))
    n2, _ = read_fastq(os.path.join(GOLD, "sample_reads_2.fastq.gz"))
    return idx_dir, rb.Index(idx_dir, 0), s1, s2, L, n1, n2


@pytest.fixture(scope="module")
def synth_small():
    idx_dir = os.path.join(GOLD, "synth_idx") + "/"
    tx = SynthTxome(META["seed"], META["genes"])
    s1, s2 = tx.reads(META["pairs"], rseed=META["rseed"], sub=META["sub"], ins=META["ins"], dele=META["del"], nn=META["n"])
    names = [f"r{i}" for i in range(META["pairs"])]
    return idx_dir, rb.Index(idx_dir, 0), s1, s2, 100, tx


def flag_opts(fname):
    return opts_from_flags(GOLDEN[f"synth/{fname}"]["flags"])


def opts_from_flags(flags):
    o = rb.default_opts()
    it = iter(flags)
    bt2 = strict = False
    for a in it:
        if a == "-s": o.sel_aln = 1
        elif a == "--hardFilter": o.hard_filter = 1
        elif a == "--mimicBT2": bt2 = True
        elif a == "--mimicStrictBT2": strict = True
        elif a == "--dpBandwidth": o.dp_bandwidth = int(next(it))
        elif a == "--ma": o.match_score = int(next(it))
        elif a == "--mm": o.mismatch_penalty = int(next(it))
        elif a == "--go": o.gap_open_penalty = int(next(it))
        elif a == "--ge": o.gap_extend_penalty = int(next(it))
        elif a == "--minScoreFrac": o.min_score_fraction = float(next(it))
        elif a == "--noOrphans": o.no_orphans = 1
        elif a == "--noDovetail": o.no_dovetail = 1
        elif a == "--consensusSlack": o.consensus_slack = float(next(it))
        elif a == "--maxMMPExtension": o.max_mmp_extension = int(next(it))
        elif a == "-f": o.fuzzy = 1
        elif a == "-m": o.max_num_hits = int(next(it))
        elif a == "-z": o.quasi_coverage = float(next(it))
        elif a == "--noSensitive": o.sensitive = 0
        elif a == "--noStrictCheck": o.strict_check = 0
        elif a == "--recoverOrphans": o.recover_orphans = 1
        else: raise ValueError(a)
    if bt2 or strict:  # reference src/RapMapSAMapper.cpp:1150-1174
        o.sel_aln = 1; o.alignment_policy = 1 if bt2 else 2; o.no_orphans = 1; o.no_dovetail = 1; o.consensus_slack = 0.35; o.max_num_hits = 1000
        if strict:
            o.min_score_fraction = 0.8; o.match_score = 1; o.mismatch_penalty = 0; o.gap_open_penalty = 25; o.gap_extend_penalty = 25
    return o


@pytest.mark.parametrize("sel", [False, True])
def test_sample_records_and_sam_match_reference(sample, sel):
    idx_dir, index, s1, s2, L, n1, n2 = sample
    opts = rb.default_opts(sel_aln=sel)
    n = s1.shape[0]
    mapper = make_mapper(index, opts, n, L)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=L)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1, s2, L), "sample")
    sam = index.sam_header() + mapper.format_sam(s1, s2, n1, n2, res, n, fixed_len=L)
    key = "sample/selaln" if sel else "sample/default"
    assert md5(sam) == GOLDEN[key]["md5"], "SAM differs from the reference's golden SAM"


@pytest.mark.parametrize("fname", sorted(k.split("/")[1] for k in GOLDEN if k.startswith("synth/")))
def test_synth_flagsets_match_oracle_and_golden_sam(synth_small, fname):
    idx_dir, index, s1, s2, L, tx = synth_small
    opts = flag_opts(fname)
    n = s1.shape[0]
    mapper = make_mapper(index, opts, n, L)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=L)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1, s2, L), fname)
    # SAM through the host formatter must equal what the reference printed (names as the generator CLI writes them)
    truth = np.zeros((n, 4), dtype=np.int64)
    from helpers import synth_lib
    import ctypes as C
    synth_lib().synth_reads(tx.h, META["rseed"], 0, n, 100, META["sub"], META["ins"], META["del"], META["n"], s1.ctypes.data, s2.ctypes.data, truth.ctypes.data)
    names1 = [f"r{i}:{truth[i,0]}:{truth[i,1]}:{truth[i,2]}/1" for i in range(n)]
    names2 = [f"r{i}:{truth[i,0]}:{truth[i,1]}:{truth[i,2]}/2" for i in range(n)]
    sam = index.sam_header() + mapper.format_sam(s1, s2, names1, names2, res, n, fixed_len=L)
    assert md5(sam) == GOLDEN[f"synth/{fname}"]["md5"], f"{fname}: SAM differs from the reference's golden SAM"


def test_sa_interval_stage_matches_oracle(synth_small):
    """Kernel 1 alone: SAIntervalHit lists == SACollector::operator() of the oracle, read by read."""
    idx_dir, index, s1, s2, L, tx = synth_small
    opts = rb.default_opts()
    n = 400
    mapper = make_mapper(index, opts, n, L)
    mapper.map_batch(s1[:n].copy(), s2[:n].copy(), n=n, fixed_len=L)
    om = OracleMapper(idx_dir, opts)
    for r in range(2 * n):
        read = (s1 if r < n else s2)[r % n].tobytes()
        assert mapper.debug_intervals(r) == om.collect(read), f"read {r}: {read!r}"


EDGE_READS = [
    "lower", "n_mid", "n30", "n31", "lead_n", "trail_n", "len30", "len31", "len32", "homopolymer", "sub1", "del2", "ins2", "iupac", "u_base", "unrelated_mate",
    "overhang_start", "overhang_end", "all_n", "short_both",
]


def edge_cases(tx_text: str):
    """Crafted pairs after SURVEY.md Appendix D, cut from a real transcript so that they map."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rc = lambda s: "".join(comp.get(c, "N") for c in reversed(s))
    frag = tx_text[40:340]
    m1, m2 = frag[:100], rc(frag[-100:])
    out = {}
    out["lower"] = (m1.lower(), m2.lower())
    out["n_mid"] = (m1[:50] + "N" + m1[51:], m2)
    out["n30"] = (m1[:30] + "N" + m1[31:], m2)
    out["n31"] = (m1[:31] + "n" + m1[32:], m2)
    out["lead_n"] = ("N" * 40 + m1[40:], m2)
    out["trail_n"] = (m1[:60] + "N" * 40, m2)
    out["len30"] = (m1[:30], m2[:30])
    out["len31"] = (m1[:31], m2[:31])
    out["len32"] = (m1[:32], m2[:32])
    out["homopolymer"] = ("A" * 100, "T" * 100)
    out["sub1"] = (m1[:45] + comp[m1[45]] + m1[46:], m2)
    out["del2"] = (m1[:60] + m1[62:] + frag[100:102], m2)
    out["ins2"] = (m1[:60] + "GT" + m1[60:98], m2)
    out["iupac"] = (m1[:20] + "R" + m1[21:], m2[:70] + "Y" + m2[71:])
    out["u_base"] = (m1.replace("T", "U", 3), m2[:50] + m2[50:].replace("T", "u", 2))
    out["unrelated_mate"] = (m1, "ACGTTGCA" * 12 + "ACGT")
    out["overhang_start"] = ("ACGTACGTAC" + tx_text[:90], rc(tx_text[150:250]))
    out["overhang_end"] = (m1, rc(frag[-90:] + "ACGTACGTAC"))
    out["all_n"] = ("N" * 100, "N" * 100)
    out["short_both"] = (m1[:10], "")
    return out


@pytest.mark.parametrize("sel", [False, True])
def test_edge_case_reads_ragged(synth_small, sel):
    """Ragged batch (offset arrays) of crafted reads: Ns, IUPAC, U, lower case, length 0/10/30/31/32, overhangs, orphans."""
    idx_dir, index, s1, s2, L, tx = synth_small
    first = read_first_transcript(idx_dir)
    cases = edge_cases(first)
    a = [cases[k][0] for k in EDGE_READS]
    b = [cases[k][1] for k in EDGE_READS]
    # plus the same pairs with mates swapped (other strand)
    a, b = a + b, b + a
    b1, o1 = pack_ragged(a)
    b2, o2 = pack_ragged(b)
    opts = rb.default_opts(sel_aln=sel)
    mapper = make_mapper(index, opts, len(a), 120)
    res = mapper.map_batch(b1, b2, n=len(a), off1=o1, off2=o2)
    ref = OracleMapper(idx_dir, opts).map(b1, b2, 0, o1, o2)
    assert_same(res, ref, "edge cases")
    assert res.num_hits > 10


def read_first_transcript(idx_dir):
    import struct
    with open(os.path.join(idx_dir, "txpInfo.bin"), "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        for _ in range(n):
            l = struct.unpack("<Q", f.read(8))[0]
            f.read(l)
        n = struct.unpack("<Q", f.read(8))[0]
        f.read(4 * n)
        n = struct.unpack("<Q", f.read(8))[0]
        text = f.read(n).decode()
    return text.split("$")[0]


def test_empty_and_single_read_batches(synth_small):
    idx_dir, index, s1, s2, L, tx = synth_small
    opts = rb.default_opts()
    mapper = make_mapper(index, opts, 8, L)
    res = mapper.map_batch(s1[:1].copy(), s2[:1].copy(), n=1, fixed_len=L)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1[:1].copy(), s2[:1].copy(), L), "n=1")
    hb = mapper.map_batch(s1[:0].copy(), s2[:0].copy(), n=0, fixed_len=L)
    assert hb.num_hits == 0


@pytest.mark.parametrize("fname", sorted(k.split("/")[1] for k in GOLDEN if k.startswith("synth_r/")))
def test_unmated_reads_match_oracle_and_golden_sam(synth_small, fname):
    """quasimap -r (processReadsSingleSA): records == oracle, SAM text == what the reference printed."""
    idx_dir, index, s1, s2, L, tx = synth_small
    opts = opts_from_flags(GOLDEN[f"synth_r/{fname}"]["flags"])
    n = s1.shape[0]
    mapper = make_mapper(index, opts, n, L)
    res = mapper.map_batch(s1, None, n=n, fixed_len=L)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1, None, L, n=n), f"unmated {fname}")
    sam = index.sam_header() + mapper.format_sam(s1, None, synth_names(tx, n)[0], None, res, n, fixed_len=L)
    assert md5(sam) == GOLDEN[f"synth_r/{fname}"]["md5"], f"unmated {fname}: SAM differs from the reference's golden SAM"


def synth_names(tx, n):
    """Read names as the generator CLI writes them into FASTQ (r<i>:<txp>:<pos>:<fraglen>/<mate>)."""
    from helpers import synth_lib

    truth = np.zeros((n, 4), dtype=np.int64)
    a, b = np.empty((n, 100), dtype=np.uint8), np.empty((n, 100), dtype=np.uint8)
    synth_lib().synth_reads(tx.h, META["rseed"], 0, n, 100, META["sub"], META["ins"], META["del"], META["n"], a.ctypes.data, b.ctypes.data, truth.ctypes.data)
    return ([f"r{i}:{truth[i,0]}:{truth[i,1]}:{truth[i,2]}/1" for i in range(n)], [f"r{i}:{truth[i,0]}:{truth[i,1]}:{truth[i,2]}/2" for i in range(n)])


def test_sam_formatting_threads_give_identical_text(synth_small):
    idx_dir, index, s1, s2, L, tx = synth_small
    opts = rb.default_opts(sel_aln=True)
    n = s1.shape[0]
    mapper = make_mapper(index, opts, n, L)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=L)
    res = rb.BatchResult(res.hits.copy(), res.pair_offsets.copy(), res.counters, res.num_hits)
    n1, n2 = synth_names(tx, n)
    one = mapper.format_sam(s1, s2, n1, n2, rb.BatchResult(res.hits.copy(), res.pair_offsets, res.counters, res.num_hits), n, fixed_len=L, threads=1)
    many = mapper.format_sam(s1, s2, n1, n2, rb.BatchResult(res.hits.copy(), res.pair_offsets, res.counters, res.num_hits), n, fixed_len=L, threads=7)
    assert one == many and md5(index.sam_header() + one) == GOLDEN["synth/selaln"]["md5"]


def test_device_resident_inputs_match_host_inputs(synth_small):
    import torch

    idx_dir, index, s1, s2, L, tx = synth_small
    opts = rb.default_opts()
    n = s1.shape[0]
    mapper = make_mapper(index, opts, n, L)
    host = mapper.map_batch(s1, s2, n=n, fixed_len=L)
    h_hits, h_off = host.hits.copy(), host.pair_offsets.copy()
    d1, d2 = torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda()
    dev = mapper.map_batch(d1, d2, n=n, fixed_len=L, location=rb.LOC_DEVICE)
    assert np.array_equal(dev.hits, h_hits) and np.array_equal(dev.pair_offsets, h_off)


PHF_MODES = ["derived_table", "walk"]   # -p lookups: dense table derived from FrugalBooMap::find at load (default) / BooPHF walked per lookup


def _phf_mode(monkeypatch, mode):
    if mode == "walk":
        monkeypatch.setenv("RAPMAP_B200_PHF", "walk")
    else:
        monkeypatch.delenv("RAPMAP_B200_PHF", raising=False)


@pytest.mark.parametrize("mode", PHF_MODES)
def test_perfect_hash_index_gives_identical_hits(synth_small, monkeypatch, mode):
    _phf_mode(monkeypatch, mode)
    idx_dir, index, s1, s2, L, tx = synth_small
    opts = rb.default_opts()
    n = s1.shape[0]
    a = make_mapper(index, opts, n, L).map_batch(s1, s2, n=n, fixed_len=L)
    a_hits, a_off = a.hits.copy(), a.pair_offsets.copy()
    pidx = rb.Index(os.path.join(GOLD, "synth_idx_p"), 0)
    b = make_mapper(pidx, opts, n, L).map_batch(s1, s2, n=n, fixed_len=L)
    assert np.array_equal(a_hits, b.hits) and np.array_equal(a_off, b.pair_offsets)


@pytest.mark.parametrize("mode", PHF_MODES)
@pytest.mark.parametrize("sel", [False, True])
def test_perfect_hash_index_matches_reference_golden_sam(synth_small, sel, monkeypatch, mode):
    """-p index (lookups through the table derived from FrugalBooMap::find, or through the on-device BooPHF walk): SAM identical to
    `rapmap_ref quasimap` on its own -p index."""
    _phf_mode(monkeypatch, mode)
    idx_dir, index, s1, s2, L, tx = synth_small
    n = s1.shape[0]
    opts = rb.default_opts(sel_aln=sel)
    pidx = rb.Index(os.path.join(GOLD, "synth_idx_p"), 0)
    mapper = make_mapper(pidx, opts, n, L)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=L)
    assert_same(res, OracleMapper(os.path.join(GOLD, "synth_idx_p") + "/", opts).map(s1, s2, L), "perfect hash")
    from helpers import synth_lib
    truth = np.zeros((n, 4), dtype=np.int64)
    a, b = s1.copy(), s2.copy()
    synth_lib().synth_reads(tx.h, META["rseed"], 0, n, 100, META["sub"], META["ins"], META["del"], META["n"], a.ctypes.data, b.ctypes.data, truth.ctypes.data)
    names1 = [f"r{i}:{truth[i,0]}:{truth[i,1]}:{truth[i,2]}/1" for i in range(n)]
    names2 = [f"r{i}:{truth[i,0]}:{truth[i,1]}:{truth[i,2]}/2" for i in range(n)]
    sam = pidx.sam_header() + mapper.format_sam(s1, s2, names1, names2, res, n, fixed_len=L)
    assert md5(sam) == GOLDEN["synth_p/selaln" if sel else "synth_p/default"]["md5"]


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the mid-size index")
@pytest.mark.parametrize("mode", PHF_MODES)
def test_mid_size_perfect_hash_index_matches_dense(monkeypatch, mode):
    """13.9k-transcript -p index (6.1 M keys over 25 BooPHF levels) gives the hits of the dense index."""
    _phf_mode(monkeypatch, mode)
    d_dir, tx = synth_index(2500)
    p_dir, _ = synth_index(2500, perfect=True)
    n = 30000
    s1, s2 = tx.reads(n, rseed=11)
    opts = rb.default_opts()
    a = make_mapper(rb.Index(d_dir, 0), opts, n, 100).map_batch(s1, s2, n=n, fixed_len=100)
    a_hits, a_off = a.hits.copy(), a.pair_offsets.copy()
    b = make_mapper(rb.Index(p_dir, 0), opts, n, 100).map_batch(s1, s2, n=n, fixed_len=100)
    assert np.array_equal(a_hits, b.hits) and np.array_equal(a_off, b.pair_offsets)


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the mid-size index")
@pytest.mark.parametrize("sel", [False, True])
def test_mid_size_synthetic_matches_oracle(sel):
    """13.9k-transcript isoform-structured index, 60k noisy pairs: records and counters identical to the oracle."""
    idx_dir, tx = synth_index(2500)
    n = 60000
    s1, s2 = tx.reads(n)
    opts = rb.default_opts(sel_aln=sel)
    index = rb.Index(idx_dir, 0)
    mapper = make_mapper(index, opts, n, 100)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=100)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1, s2, 100), f"mid sel={sel}")
    # size-independent properties: offsets monotone, every pair's hits sorted by tid for paired hits, idempotence
    off = res.pair_offsets.astype(np.int64)
    assert (np.diff(off) >= 0).all() and off[-1] == res.num_hits
    again = mapper.map_batch(s1, s2, n=n, fixed_len=100)
    assert np.array_equal(again.hits, res.hits)


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the repeat-family index")
def test_repeat_families_exercise_big_intervals_and_spill():
    """Transcriptome with repeat families (intervals near the 1000 cap, long hit lists): global work-strip path."""
    idx_dir, tx = synth_index(600, seed=4711, repeats=6)
    n = 20000
    s1, s2 = tx.reads(n, rseed=7)
    opts = rb.default_opts()
    index = rb.Index(idx_dir, 0)
    mapper = make_mapper(index, opts, n, 100)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=100)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1, s2, 100), "repeats")


@pytest.mark.parametrize("read_len,sel,max_len", [(150, False, 0), (150, True, 0), (250, True, 0), (300, False, 0), (300, True, 1000), (700, False, 1000)])
def test_longer_reads_match_oracle(read_len, sel, max_len):
    """Reads longer than the benchmark's 100 bases: 5..10 packed words per read in the SA-lookup kernel, ksw2 windows
    beyond the thread-per-job kernel's strip (general warp kernel), more intervals / SA entries per read in hit resolution."""
    idx_dir, tx = synth_index(2500)
    n = 6000
    s1, s2 = tx.reads(n, rseed=4242, read_len=read_len)
    opts = rb.default_opts(sel_aln=sel)
    index = rb.Index(idx_dir, 0)
    # max_read_len = 1000: packed words + k-mer masks no longer fit a block's shared memory, the masks are read from global memory
    mapper = make_mapper(index, opts, n, max_len or read_len)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=read_len)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1, s2, read_len), f"L={read_len} sel={sel}")


@pytest.mark.parametrize("flags", [["-s", "--recoverOrphans"], ["-f", "--recoverOrphans", "--noDovetail"], ["-s", "--recoverOrphans", "--hardFilter", "--noOrphans"]])
def test_orphan_recovery_matches_oracle(flags):
    """--recoverOrphans on 20k noisy pairs of the 13.9k-transcript index (about one pair in eight goes through recovery)."""
    idx_dir, tx = synth_index(2500)
    n = 20000
    s1, s2 = tx.reads(n, rseed=2718, sub=30000, ins=3000, dele=3000, nn=3000)
    opts = opts_from_flags(flags)
    index = rb.Index(idx_dir, 0)
    mapper = make_mapper(index, opts, n, 100)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=100)
    assert_same(res, OracleMapper(idx_dir, opts).map(s1, s2, 100), " ".join(flags))


# ---------------------------------------------------------------------------------------------------------------
# flags that only matter when a read hits BOTH strands: transcriptome with diverged antisense copies
# ---------------------------------------------------------------------------------------------------------------
DISCRIMINATING = [["--noSensitive"], ["--noStrictCheck"], ["--noSensitive", "--noStrictCheck"], ["-s"], ["-s", "--noSensitive"], ["-s", "--noStrictCheck"],
                  ["-s", "--noSensitive", "--noStrictCheck"], ["-z", "0.7"], ["-m", "2"], ["-f"], ["-f", "--noStrictCheck"]]


@pytest.fixture(scope="module")
def antisense():
    if not have_ref():
        pytest.skip("needs oracle/_ref to build the antisense index")
    from helpers import antisense_index

    idx_dir, tx, _ = antisense_index()
    n = 12000
    s1, s2 = tx.reads(n, rseed=77, sub=15000, ins=1000, dele=1000, nn=1000)
    base = OracleMapper(idx_dir, rb.default_opts()).map(s1, s2, 100)
    return idx_dir, rb.Index(idx_dir, 0), s1, s2, n, base


@pytest.mark.parametrize("flags", [[]] + DISCRIMINATING, ids=lambda f: " ".join(f) or "default")
def test_strand_and_skip_flags_on_antisense_transcriptome(antisense, flags):
    """Each flag set gives records identical to the oracle AND different from the default flags' (the fixture discriminates:
    a wrong NIP skip, lce, k-mer vote or strand filter cannot hide)."""
    idx_dir, index, s1, s2, n, base = antisense
    opts = opts_from_flags(flags)
    ref = OracleMapper(idx_dir, opts).map(s1, s2, 100)
    if flags:
        assert not (ref.num_hits == base.num_hits and np.array_equal(ref.hits, base.hits)), f"{flags}: oracle output equals the default output"
    mapper = make_mapper(index, opts, n, 100)
    assert_same(mapper.map_batch(s1, s2, n=n, fixed_len=100), ref, " ".join(flags))


@pytest.mark.parametrize("flags", [[], ["--noSensitive"], ["--noStrictCheck"], ["--noSensitive", "--noStrictCheck"], ["-s"], ["-s", "--noSensitive"]],
                         ids=lambda f: " ".join(f) or "default")
def test_sa_interval_stage_under_flags_on_antisense_transcriptome(antisense, flags):
    """Kernel 1 alone under the flags that steer it: SAIntervalHit lists == SACollector::operator() of the oracle, read by read;
    under every non-default flag set at least one read's lists differ from the default ones."""
    idx_dir, index, s1, s2, _, _ = antisense
    n = 1500
    opts = opts_from_flags(flags)
    mapper = make_mapper(index, opts, n, 100)
    mapper.map_batch(s1[:n].copy(), s2[:n].copy(), n=n, fixed_len=100)
    om, om0 = OracleMapper(idx_dir, opts), OracleMapper(idx_dir, rb.default_opts())
    differs = 0
    for r in range(2 * n):
        read = (s1 if r < n else s2)[r % n].tobytes()
        exp = om.collect(read)
        assert mapper.debug_intervals(r) == exp, f"read {r}: {read!r}"
        differs += exp != om0.collect(read)
    if flags:
        assert differs > 0, f"{flags}: no read's interval lists differ from the default walk"


# ---------------------------------------------------------------------------------------------------------------
# the boundary: overflow -> grow -> re-run, asynchronous calls, concurrent mappers, index replicas
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the mid-size index")
@pytest.mark.parametrize("flags", [[], ["-s"], ["-f"], ["-s", "--recoverOrphans"]], ids=lambda f: " ".join(f) or "default")
def test_every_arena_overflow_retry_path(monkeypatch, flags):
    """RAPMAP_B200_TINY_ARENAS=1 starts every growable device work area far too small (interval arena 64 records, 1 interval
    per strand in the per-thread lists, 8 hit records, 8 QA records, 16 positions, 128-entry work strips): the first batch
    must walk through each overflow -> grow -> re-run path and still equal the oracle; the second batch runs clean."""
    monkeypatch.setenv("RAPMAP_B200_TINY_ARENAS", "1")
    idx_dir, tx = synth_index(600, seed=4711, repeats=6)
    n = 8000
    s1, s2 = tx.reads(n, rseed=7)
    opts = opts_from_flags(flags)
    mapper = make_mapper(rb.Index(idx_dir, 0), opts, n, 100)
    monkeypatch.delenv("RAPMAP_B200_TINY_ARENAS")
    ref = OracleMapper(idx_dir, opts).map(s1, s2, 100)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=100)
    assert mapper.timing().retries >= 2, "the tiny arenas did not force a re-run"
    assert_same(res, ref, "after forced retries")
    res2 = mapper.map_batch(s1, s2, n=n, fixed_len=100)
    assert mapper.timing().retries == 0
    assert_same(res2, ref, "second batch")


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the mid-size index")
@pytest.mark.parametrize("sel,pinned", [(False, False), (True, False), (False, True), (True, True)])
def test_async_pipeline_chunks_in_flight(sel, pinned):
    """rapmap_cuda_map_batch_async / rapmap_cuda_mapper_wait: one host thread, two mappers on one shared index, each with several
    chunks in flight (copy-in, kernels and copy-out of consecutive chunks overlap on the mapper's three streams); every
    chunk equals the oracle's result for that chunk.  Pageable and pinned output buffers; a third chunk is refused."""
    import torch

    idx_dir, tx = synth_index(2500)
    index = rb.Index(idx_dir, 0)
    opts = rb.default_opts(sel_aln=sel)
    n, chunks = 5000, 14
    data = [tx.reads(n, rseed=1000 + c) for c in range(chunks)]
    om = OracleMapper(idx_dir, opts)
    refs = [om.map(a, b, 100) for a, b in data]
    mappers = [make_mapper(index, opts, n, 100) for _ in range(2)]

    def buffers():
        if pinned:
            return torch.empty(16 * n * 28, dtype=torch.uint8).pin_memory(), torch.empty(n + 1, dtype=torch.int64).pin_memory()
        return np.empty(16 * n, dtype=rb.HIT_DTYPE), np.empty(n + 1, dtype=np.uint64)

    D = rb.max_in_flight()
    outs = [[buffers() for _ in range(D)] for _ in range(2)]
    done = {}
    order = [[], []]

    def collect(k):
        c = order[k].pop(0)
        r = mappers[k].wait()
        if pinned:
            hits = r.hits.numpy()[: r.num_hits * 28].copy().view(rb.HIT_DTYPE)
            offs = r.pair_offsets.numpy().copy().view(np.uint64)
        else:
            hits, offs = r.hits.copy(), r.pair_offsets.copy()
        done[c] = rb.BatchResult(hits, offs, r.counters, r.num_hits)

    for c in range(chunks):
        k = c % 2
        if mappers[k].in_flight == D:
            collect(k)
        ho, oo = outs[k][(c // 2) % D]
        mappers[k].map_batch_async(data[c][0], data[c][1], n=n, fixed_len=100, hits_out=ho, offsets_out=oo, capacity=16 * n)
        order[k].append(c)
    while mappers[0].in_flight < D:  # fill the pipeline, then one chunk more is refused
        c = chunks - 1
        ho, oo = buffers()
        mappers[0].map_batch_async(data[c][0], data[c][1], n=n, fixed_len=100, hits_out=ho, offsets_out=oo, capacity=16 * n)
        order[0].append(c)
    with pytest.raises(rb.RapMapCudaError):
        ho, oo = buffers()
        mappers[0].map_batch_async(data[0][0], data[0][1], n=n, fixed_len=100, hits_out=ho, offsets_out=oo, capacity=16 * n)
    for k in range(2):
        while mappers[k].in_flight:
            collect(k)
    for c in range(chunks):
        assert_same(done[c], refs[c], f"chunk {c}")
    with pytest.raises(rb.RapMapCudaError):
        mappers[0].wait()  # nothing in flight


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the mid-size index")
@pytest.mark.parametrize("sel", [False, True])
def test_copy_engine_copy_out_with_uneven_chunks(sel, monkeypatch):
    """Pinned output buffers: the copy engines move the offsets and the part of the records the previous chunks predict,
    copy_out_kernel only the tail.  Chunks of very different size and hit density (unmappable reads: the prediction is far
    above the real count; a dense chunk after them: far below) must still arrive exactly."""
    import torch

    monkeypatch.setenv("RAPMAP_B200_COPYOUT", "engine")  # the copy-engine part for these small chunks too (default: >= 32k pairs)
    idx_dir, tx = synth_index(2500)
    index = rb.Index(idx_dir, 0)
    opts = rb.default_opts(sel_aln=sel)
    rng = np.random.default_rng(5)
    junk = lambda n: (np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, 100))], np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, 100))])
    plan = [("reads", 4000), ("reads", 4000), ("junk", 4000), ("reads", 1500), ("junk", 300), ("reads", 4000), ("reads", 37)]
    om = OracleMapper(idx_dir, opts)
    mapper = make_mapper(index, opts, 4000, 100)
    ho, oo = torch.empty(16 * 4000 * 28, dtype=torch.uint8).pin_memory(), torch.empty(4001, dtype=torch.int64).pin_memory()
    for c, (kind, n) in enumerate(plan):
        a, b = tx.reads(n, rseed=3000 + c) if kind == "reads" else junk(n)
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        ho.fill_(0xEE)
        mapper.map_batch_async(a, b, n=n, fixed_len=100, hits_out=ho, offsets_out=oo, capacity=16 * 4000)
        r = mapper.wait()
        got = rb.BatchResult(r.hits.numpy()[: r.num_hits * 28].copy().view(rb.HIT_DTYPE), r.pair_offsets.numpy()[: n + 1].copy().view(np.uint64), r.counters, r.num_hits)
        assert_same(got, om.map(a, b, 100), f"chunk {c} ({kind}, {n} pairs)")


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the mid-size index")
def test_overflow_with_two_chunks_in_flight(monkeypatch):
    """Tiny arenas AND several chunks in flight: the first chunk's overflow re-allocates the work areas under the second chunk,
    whose attempt is then repeated; both must equal the oracle."""
    monkeypatch.setenv("RAPMAP_B200_TINY_ARENAS", "1")
    idx_dir, tx = synth_index(2500)
    index = rb.Index(idx_dir, 0)
    opts = rb.default_opts(sel_aln=True)
    n = 6000
    data = [tx.reads(n, rseed=77 + c) for c in range(3)]
    mapper = make_mapper(index, opts, n, 100)
    monkeypatch.delenv("RAPMAP_B200_TINY_ARENAS")
    om = OracleMapper(idx_dir, opts)
    outs = [(np.empty(16 * n, dtype=rb.HIT_DTYPE), np.empty(n + 1, dtype=np.uint64)) for _ in range(3)]
    got = []
    for c in range(3):
        if mapper.in_flight == rb.max_in_flight():
            r = mapper.wait()
            got.append(rb.BatchResult(r.hits.copy(), r.pair_offsets.copy(), r.counters, r.num_hits))
        mapper.map_batch_async(data[c][0], data[c][1], n=n, fixed_len=100, hits_out=outs[c][0], offsets_out=outs[c][1], capacity=16 * n)
    while mapper.in_flight:
        r = mapper.wait()
        got.append(rb.BatchResult(r.hits.copy(), r.pair_offsets.copy(), r.counters, r.num_hits))
    for c in range(3):
        assert_same(got[c], om.map(data[c][0], data[c][1], 100), f"chunk {c}")


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref to build the mid-size index")
def test_concurrent_mappers_on_host_threads():
    """Four host threads, each with its own mapper on ONE shared index (the reference's worker-thread model): every chunk of
    every thread equals the oracle."""
    import threading

    idx_dir, tx = synth_index(2500)
    index = rb.Index(idx_dir, 0)
    opts = rb.default_opts(sel_aln=True)
    n, per = 4000, 5
    data = [[tx.reads(n, rseed=5000 + 10 * t + c) for c in range(per)] for t in range(4)]
    om = OracleMapper(idx_dir, opts)
    refs = [[om.map(a, b, 100) for a, b in row] for row in data]
    got = [[None] * per for _ in range(4)]
    errs = []

    def work(t):
        try:
            mp = rb.Mapper(index, opts, max_batch=n, max_read_len=100)
            for c in range(per):
                r = mp.map_batch(data[t][c][0], data[t][c][1], n=n, fixed_len=100)
                got[t][c] = rb.BatchResult(r.hits.copy(), r.pair_offsets.copy(), r.counters, r.num_hits)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [th.start() for th in ths]
    [th.join() for th in ths]
    assert not errs, errs
    for t in range(4):
        for c in range(per):
            assert_same(got[t][c], refs[t][c], f"thread {t} chunk {c}")


def test_index_replica_from_image_carries_names(synth_small):
    """An index built around a copy of the packed image (what another rank receives over NCCL) maps identically and can print
    SAM: transcript names and lengths travel inside the image."""
    import torch

    idx_dir, index, s1, s2, L, tx = synth_small
    ptr, nbytes = index.image()
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    src = torch.empty(0, dtype=torch.uint8, device="cuda")
    # device-to-device copy of the image through the CUDA runtime
    import ctypes as C
    try:
        rt = C.CDLL("libcudart.so.12")
    except OSError:
        rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
    assert rt.cudaMemcpy(C.c_void_p(blob.data_ptr()), C.c_void_p(ptr), C.c_size_t(nbytes), 3) == 0
    replica = rb.Index.from_image(0, blob.data_ptr(), nbytes)
    assert replica.sam_header() == index.sam_header()
    assert replica.transcript_name(3) == index.transcript_name(3) and replica.transcript_len(3) == index.transcript_len(3)
    opts = rb.default_opts()
    n = s1.shape[0]
    a = make_mapper(index, opts, n, L).map_batch(s1, s2, n=n, fixed_len=L)
    a_hits = a.hits.copy()
    b = make_mapper(replica, opts, n, L).map_batch(s1, s2, n=n, fixed_len=L)
    assert np.array_equal(a_hits, b.hits)
    with pytest.raises(rb.RapMapCudaError):
        rb.Index.from_image(0, blob.data_ptr() + 32, nbytes - 32)  # misaligned / not an image
    del src


def test_perfect_hash_index_replica_from_image(synth_small):
    """The image of a -p index (BooPHF arrays, the table derived from them, the transcript names behind them) as a replica:
    same SAM header, same hits as the dense index."""
    import torch

    idx_dir, index, s1, s2, L, tx = synth_small
    pidx = rb.Index(os.path.join(GOLD, "synth_idx_p"), 0)
    ptr, nbytes = pidx.image()
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    import ctypes as C
    try:
        rt = C.CDLL("libcudart.so.12")
    except OSError:
        rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
    assert rt.cudaMemcpy(C.c_void_p(blob.data_ptr()), C.c_void_p(ptr), C.c_size_t(nbytes), 3) == 0
    replica = rb.Index.from_image(0, blob.data_ptr(), nbytes)
    assert replica.sam_header() == pidx.sam_header()
    opts = rb.default_opts()
    n = s1.shape[0]
    a = make_mapper(index, opts, n, L).map_batch(s1, s2, n=n, fixed_len=L)
    a_hits, a_off = a.hits.copy(), a.pair_offsets.copy()
    b = make_mapper(replica, opts, n, L).map_batch(s1, s2, n=n, fixed_len=L)
    assert np.array_equal(a_hits, b.hits) and np.array_equal(a_off, b.pair_offsets)


@pytest.mark.parametrize("sel", [False, True])
def test_direct_copy_out_to_pinned_and_device_buffers(synth_small, sel):
    """Results leave the device by a kernel when the caller's buffers are device memory or pinned host memory (one host
    synchronisation per batch); pageable numpy buffers take the cudaMemcpyAsync path.  All three must hold the same bytes;
    a too small pinned buffer reports RAPMAP_ERR_CAPACITY with the needed count."""
    import torch

    idx_dir, index, s1, s2, L, tx = synth_small
    opts = rb.default_opts(sel_aln=sel)
    n = s1.shape[0]
    mapper = make_mapper(index, opts, n, L)
    ref = mapper.map_batch(s1, s2, n=n, fixed_len=L)
    ref_hits, ref_off, nh = ref.hits.copy().view(np.uint8).reshape(-1), ref.pair_offsets.copy(), ref.num_hits
    cap = nh + 7
    for where in ("pinned", "device"):
        hits = torch.zeros(cap * 28, dtype=torch.uint8)
        offs = torch.zeros(n + 1, dtype=torch.int64)
        hits, offs = (hits.pin_memory(), offs.pin_memory()) if where == "pinned" else (hits.cuda(), offs.cuda())
        r = mapper.map_batch(s1, s2, n=n, fixed_len=L, hits_out=hits, offsets_out=offs, out_location=rb.LOC_HOST if where == "pinned" else rb.LOC_DEVICE, capacity=cap)
        torch.cuda.synchronize()
        assert r.num_hits == nh and mapper.timing().retries == 0
        assert np.array_equal(hits.cpu().numpy()[: nh * 28], ref_hits), where
        assert np.array_equal(offs.cpu().numpy().view(np.uint64), ref_off), where
    small = torch.zeros(10 * 28, dtype=torch.uint8).pin_memory()
    offs = torch.zeros(n + 1, dtype=torch.int64).pin_memory()
    with pytest.raises(rb.RapMapCudaError) as e:
        mapper.map_batch(s1, s2, n=n, fixed_len=L, hits_out=small, offsets_out=offs, capacity=10)
    assert e.value.code == rb.ERR_CAPACITY


@pytest.mark.parametrize("idx,case", [("synth_idx", "synth/default"), ("synth_idx", "synth/selaln"), ("synth_idx_p", "synth_p/selaln")])
def test_bigsa_index_flavour(synth_small, tmp_path, idx, case):
    """BigSA (int64) index files (helpers.make_bigsa_copy; accepted by the unmodified reference, tests/test_oracle_golden.py):
    read, narrowed to the device's 32-bit positions, and the SAM is the golden SAM of the 32-bit index."""
    from helpers import make_bigsa_copy

    idx_dir, index, s1, s2, L, tx = synth_small
    big = rb.Index(make_bigsa_copy(os.path.join(GOLD, idx), str(tmp_path / "big")), 0)
    opts = opts_from_flags(GOLDEN[case]["flags"])
    n = s1.shape[0]
    mapper = make_mapper(big, opts, n, L)
    res = mapper.map_batch(s1, s2, n=n, fixed_len=L)
    n1, n2 = synth_names(tx, n)
    sam = big.sam_header() + mapper.format_sam(s1, s2, n1, n2, res, n, fixed_len=L)
    assert md5(sam) == GOLDEN[case]["md5"]
