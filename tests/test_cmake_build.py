"""CPU: the CMake package (CMakeLists.txt: C++14 host, CUDA sm_100a) builds librapmap_cuda.so and the CLI, installs, and a
separate consumer project finds it with find_package(rapmap_b200) and links rapmap_b200::rapmap_cuda - the way the
reference's own src/CMakeLists.txt would pull it in."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

import rapmap_b200 as rb
from helpers import ROOT

CMAKE = shutil.which("cmake")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(CMAKE is None or not os.path.exists(NVCC), reason="cmake / nvcc not available")
def test_cmake_build_install_and_find_package(tmp_path):
    bld, pre = tmp_path / "build", tmp_path / "prefix"
    run = lambda cmd, **kw: subprocess.run(cmd, check=True, capture_output=True, text=True, **kw)
    run([CMAKE, "-S", ROOT, "-B", str(bld), f"-DCMAKE_CUDA_COMPILER={NVCC}", "-DCMAKE_BUILD_TYPE=Release"])
    run([CMAKE, "--build", str(bld), "-j", str(min(8, os.cpu_count() or 1))])
    run([CMAKE, "--install", str(bld), "--prefix", str(pre)])
    so = pre / "lib" / "librapmap_cuda.so"
    assert so.exists() and (pre / "bin" / "rapmap_b200").exists()
    assert (pre / "include" / "rapmap_cuda.h").exists() and (pre / "include" / "rapmap_b200" / "adapter.hpp").exists()
    L = C.CDLL(str(so))
    for s in rb.SYMBOLS:
        assert hasattr(L, s), s
    elfs = run(["cuobjdump", "--list-elf", str(so)]).stdout
    assert "sm_100a" in elfs and "sm_90" not in elfs and "sm_80" not in elfs, elfs
    # a consumer project
    cons = tmp_path / "consumer"
    cons.mkdir()
    (cons / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.24)\nproject(consumer LANGUAGES CXX)\nset(CMAKE_CXX_STANDARD 14)\n"
        "find_package(rapmap_b200 REQUIRED)\nadd_executable(consumer main.cpp)\ntarget_link_libraries(consumer PRIVATE rapmap_b200::rapmap_cuda)\n")
    (cons / "main.cpp").write_text(
        '#include "rapmap_cuda.h"\n#include <cstdio>\nint main() { rapmap_cuda_opts_t o; rapmap_cuda_opts_selaln(&o);\n'
        '  rapmap_cuda_index_t* idx = nullptr; int rc = rapmap_cuda_index_load("/nonexistent/", 0, &idx);\n'
        '  std::printf("%u %d %d %s\\n", o.max_num_hits, (int)o.sel_aln, rc, rapmap_cuda_last_error()); return 0; }\n')
    run([CMAKE, "-S", str(cons), "-B", str(cons / "b"), f"-DCMAKE_PREFIX_PATH={pre}"])
    run([CMAKE, "--build", str(cons / "b")])
    out = run([str(cons / "b" / "consumer")], env=dict(os.environ, LD_LIBRARY_PATH=str(pre / "lib"))).stdout.split()
    assert out[0] == "200" and out[1] == "1" and out[2] == str(rb.ERR_IO), out
