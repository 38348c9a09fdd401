"""Builds the C++ host library (libdomain_decomp.so: Grid, Partitioner, DomainUtils,
CudaRcbPartitioner over the C ABI), the `decomp` CLI and the host test executable."""
from __future__ import annotations

import os
import shutil
import subprocess

from . import build as b

HOST_LIB = os.path.join(b.PKG, "libdomain_decomp.so")
DECOMP = os.path.join(b.PKG, "decomp")
HOST_TESTS = os.path.join(b.PKG, "host_tests")
NC_TOOL = os.path.join(b.PKG, "nc_tool")

LIB_SRCS = ["PluginBench.cpp", "HostBuffer.cpp", "CdlIO.cpp", "NcClassic.cpp", "DomainUtils.cpp", "Grid.cpp", "Partitioner.cpp", "CudaRcbPartitioner.cpp"]


def _cxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def build_host(force: bool = False, verbose: bool = False) -> None:
    inc = ["-I", os.path.join(b.INCLUDE, "domain_decomp"), "-I", b.INCLUDE, "-I", b.HOST]
    flags = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra", "-pedantic", "-pthread"]
    srcs = [os.path.join(b.HOST, s) for s in LIB_SRCS]
    hdrs = [os.path.join(b.INCLUDE, "domain_decomp", h) for h in os.listdir(os.path.join(b.INCLUDE, "domain_decomp")) if h.endswith(".hpp")]
    hdrs += [os.path.join(b.HOST, "CdlIO.hpp"), os.path.join(b.HOST, "NcClassic.hpp"), os.path.join(b.INCLUDE, "ddc.h"), __file__]
    rpath = "-Wl,-rpath,$ORIGIN"

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)

    if force or b._stale(HOST_LIB, srcs + hdrs + [b.CUDA_LIB]):
        run([_cxx()] + flags + inc + ["-shared", "-o", HOST_LIB] + srcs + ["-L", b.PKG, "-lddc_cuda", rpath])
    for exe, src in ((DECOMP, "main.cpp"), (HOST_TESTS, "host_tests.cpp"), (NC_TOOL, "nc_tool.cpp")):
        s = os.path.join(b.HOST, src)
        # nc_tool also reaches the (non-exported) file readers / writers: it compiles them in
        extra = [os.path.join(b.HOST, x) for x in ("NcClassic.cpp", "CdlIO.cpp")] if exe == NC_TOOL else []
        if force or b._stale(exe, [s, HOST_LIB] + hdrs + extra):
            run([_cxx()] + flags + inc + ["-o", exe, s] + extra + ["-L", b.PKG, "-ldomain_decomp", "-lddc_cuda", rpath])
