"""The CMake build (CMakeLists.txt: the reference's package / target names, CXX + CUDA for sm_100a) configures and
builds here -- nvcc cross-compiles without a GPU -- and produces the same targets the reference's CMakeLists.txt
does (CMakeLists.txt:69-114 there): libdomain_decomp.so (+ libddc_cuda.so behind the C ABI) and `decomp`."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_cmake_configures_and_builds(tmp_path):
    cmake, nvcc = shutil.which("cmake"), shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not cmake or not os.path.exists(nvcc):
        pytest.skip("cmake / nvcc not available")
    b = str(tmp_path / "build")
    gen = ["-G", "Ninja"] if shutil.which("ninja") else []
    out = subprocess.run([cmake, "-S", ROOT, "-B", b, "-DCMAKE_CUDA_COMPILER=" + nvcc] + gen, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    out = subprocess.run([cmake, "--build", b, "--target", "decomp", "-j", "8"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    for f in ("libddc_cuda.so", "libdomain_decomp.so", "decomp"):
        assert os.path.exists(os.path.join(b, f)), f
    elf = subprocess.run(["cuobjdump", "-lelf", os.path.join(b, "libddc_cuda.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf, elf
    help_text = subprocess.run([os.path.join(b, "decomp"), "--help"], capture_output=True, text=True).stdout
    assert "--grid" in help_text and "--gpus" in help_text
