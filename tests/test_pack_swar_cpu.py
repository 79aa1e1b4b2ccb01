"""Host check of the four-bases-per-operation read packer (rapmap_b200/csrc/pack_swar.cuh is host-callable; pack_reads_kernel uses
it for full, aligned 32-base words) against the per-base rule: every byte value at every position, plus random words; and the branch-free ksw2 base code of the
pair kernel's strip fill against the switch-based nt4 / rcChar tables for every byte value."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swar_packer_equals_per_base_rule(tmp_path):
    exe = tmp_path / "pack_swar_host"
    subprocess.run(["g++", "-O2", "-std=c++14", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "pack_swar_host.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "base codes: mismatches 0" in out.stdout and "words" in out.stdout and out.stdout.strip().endswith("mismatches 0")
