"""Host-side checks of the fused kernels' data layouts (strip-tiled planes, de-interleaved match planes, the
strip x band x chunk plan, the integer cost lattice): tests/layout_check.cu, compiled with nvcc and run on the CPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layout_arithmetic_on_the_host(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "layout_check")
    lib = os.path.join(ROOT, "stereo_matching_cuda_b200", "libstereo_b200.so")
    assert os.path.exists(lib), "build the library first (python -m stereo_matching_cuda_b200.build)"
    cmd = [nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
           os.path.join(ROOT, "tests", "layout_check.cu"), "-o", exe, lib, "-Xlinker", "-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "layout_check: ok" in r.stdout, r.stdout + r.stderr
