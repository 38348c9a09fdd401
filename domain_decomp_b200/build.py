"""In-tree build of the CUDA library (and, later, the C++ host library / CLI).

    python -m domain_decomp_b200.build          # build everything that is stale
    python -m domain_decomp_b200.build --force

Everything is compiled for sm_100a only (B200); the products are git-ignored .so files next to
this package so that they travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
INCLUDE = os.path.join(ROOT, "include")

CUDA_LIB = os.path.join(PKG, "libddc_cuda.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # Zoltan's double arithmetic must not be contracted into FMAs
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, "ddc_api.cu")]
    deps = srcs + [os.path.join(CSRC, "ddc_kernels.cuh"), os.path.join(CSRC, "ddc_median.cuh"), os.path.join(CSRC, "ddc_neighbours.cuh"),
                   os.path.join(INCLUDE, "ddc.h"), __file__]
    if force or _stale(CUDA_LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-I", INCLUDE, "-I", CSRC, "-o", CUDA_LIB] + srcs + ["-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return CUDA_LIB


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda_lib(force, verbose)
    try:
        from . import build_host
    except ImportError:
        return
    build_host.build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", CUDA_LIB)
