"""CPU: the C-ABI library builds, loads, exports every symbol of include/rapmap_cuda.h, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

import rapmap_b200 as rb
from helpers import GOLD, ROOT


def test_library_exports_every_declared_symbol():
    L = rb.lib()
    hdr = open(os.path.join(ROOT, "include", "rapmap_cuda.h")).read()
    declared = set(re.findall(r"\b(rapmap_cuda_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(rb.SYMBOLS), declared ^ set(rb.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_struct_layouts_match_header():
    assert C.sizeof(rb.Hit) == 28
    assert C.sizeof(rb.SAInterval) == 32
    assert C.sizeof(rb.ReadBatch) == 48
    assert C.sizeof(rb.HitBatch) == 80
    o = rb.default_opts()
    assert (o.max_num_hits, o.sensitive, o.strict_check, o.sel_aln, o.dp_bandwidth, o.max_mmp_extension) == (200, 1, 1, 0, 15, 7)
    assert (o.match_score, o.mismatch_penalty, o.gap_open_penalty, o.gap_extend_penalty) == (2, -4, 4, 2)
    assert abs(o.consensus_slack - 0.2) < 1e-6 and o.min_score_fraction == 0.65
    assert rb.default_opts(sel_aln=True).sel_aln == 1


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rb.RapMapCudaError) as e:
        rb.Index(os.path.join(GOLD, "sample_idx"), 0)
    assert e.value.code == rb.ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_product_does_not_reference_oracle():
    """Nothing under rapmap_b200/ may import, link or execute oracle/."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "rapmap_b200")):
        if "_build" in dp or "__pycache__" in dp:
            continue
        for f in fs:
            txt = open(os.path.join(dp, f), errors="replace").read()
            if re.search(r"oracle/|liboracle|quasimap_oracle|oracle_", txt):
                bad.append(os.path.join(dp, f))
    assert not bad, bad


def _corrupt_copy(tmp_path, name, mutate):
    import shutil

    d = tmp_path / name
    shutil.copytree(os.path.join(GOLD, "sample_idx"), d)
    mutate(d)
    return str(d)


@pytest.mark.parametrize("what", ["truncated_sa", "huge_count", "sa_out_of_range", "bad_offsets", "truncated_hash", "missing_file"])
def test_malformed_index_is_an_io_error_not_a_crash(tmp_path, what):
    """rapmap_cuda_index_load on a damaged index directory returns RAPMAP_ERR_IO (no exception or abort crosses the C-ABI,
    no unbounded allocation from a corrupt size field); the files are validated before any device work."""
    import struct

    def mutate(d):
        if what == "truncated_sa":
            b = (d / "sa.bin").read_bytes()
            (d / "sa.bin").write_bytes(b[: len(b) // 2])
        elif what == "huge_count":
            b = bytearray((d / "txpInfo.bin").read_bytes())
            b[0:8] = struct.pack("<Q", 1 << 60)
            (d / "txpInfo.bin").write_bytes(bytes(b))
        elif what == "sa_out_of_range":
            b = bytearray((d / "sa.bin").read_bytes())
            b[8 + 40 : 8 + 44] = struct.pack("<i", 0x7FFFFFF0)
            (d / "sa.bin").write_bytes(bytes(b))
        elif what == "bad_offsets":
            b = bytearray((d / "txpInfo.bin").read_bytes())
            n = struct.unpack_from("<Q", b, 0)[0]
            p = 8
            for _ in range(n):
                p += 8 + struct.unpack_from("<Q", b, p)[0]
            b[p + 8 + 4 : p + 8 + 8] = struct.pack("<i", -5)  # second transcript offset
            (d / "txpInfo.bin").write_bytes(bytes(b))
        elif what == "truncated_hash":
            b = (d / "hash.bin").read_bytes()
            (d / "hash.bin").write_bytes(b[: len(b) - 100])
        elif what == "missing_file":
            os.remove(d / "rsd.bin")

    with pytest.raises(rb.RapMapCudaError) as e:
        rb.Index(_corrupt_copy(tmp_path, what, mutate), 0)
    assert e.value.code == rb.ERR_IO, str(e.value)


def test_bigsa_index_parses_and_oversized_is_refused(tmp_path):
    """A BigSA (int64) index directory is parsed and narrowed (no GPU here: the call then stops at the device check with
    RAPMAP_ERR_CUDA, not at the files); one whose transcript offsets pass 2^31 is refused with RAPMAP_ERR_UNSUPPORTED."""
    import struct

    import torch

    from helpers import make_bigsa_copy

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the GPU tests")
    big = make_bigsa_copy(os.path.join(GOLD, "synth_idx"), str(tmp_path / "big"))
    with pytest.raises(rb.RapMapCudaError) as e:
        rb.Index(big, 0)
    assert e.value.code == rb.ERR_CUDA, str(e.value)
    b = bytearray(open(os.path.join(big, "txpInfo.bin"), "rb").read())
    n = struct.unpack_from("<Q", b, 0)[0]
    p = 8
    for _ in range(n):
        p += 8 + struct.unpack_from("<Q", b, p)[0]
    struct.pack_into("<q", b, p + 8 + 8 * (n - 1), 1 << 33)  # last transcript starts beyond 2^31
    open(os.path.join(big, "txpInfo.bin"), "wb").write(bytes(b))
    with pytest.raises(rb.RapMapCudaError) as e:
        rb.Index(big, 0)
    assert e.value.code == rb.ERR_UNSUPPORTED, str(e.value)
