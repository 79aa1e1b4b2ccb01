"""Host check of the bit-vector edit distance behind --recoverOrphans (rapmap_b200/csrc/orphan_recovery.cuh is
host-callable): 40,000 random (read, window) cases against the plain semi-global DP, which is what the oracle runs and
what is pinned to the reference's edlib call through the --recoverOrphans golden SAM."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_myers_bit_vector_equals_plain_dp(tmp_path):
    exe = tmp_path / "myers_vs_dp"
    subprocess.run(["g++", "-O2", "-std=c++14", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "myers_vs_dp.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
