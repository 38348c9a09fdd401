"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, fails loudly
without a GPU (no fallback), and its pure-host helpers behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def capi():
    from domain_decomp_b200 import build, capi
    build.build_cuda_lib()
    return capi


def test_header_symbols_are_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "ddc.h")).read()
    declared = set(re.findall(r"DDC_API\s+[\w\s\*]+?\b(ddc_\w+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = capi.load()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.ddc_version()


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DdcError, match="no CUDA device|CUDA"):
        capi.Handle(0)


def test_shard_rows(capi):
    for ny, g in ((32768, 8), (522, 4), (5, 8), (7, 2), (1, 1)):
        rows = [capi.shard_rows(ny, g, r) for r in range(g)]
        assert rows[0][0] == 0
        assert sum(c for _, c in rows) == ny
        rpr = -(-ny // g)
        for r, (b, c) in enumerate(rows):
            assert b == min(ny, r * rpr) and 0 <= c <= rpr


def test_synthetic_mask_host_is_deterministic_and_sharded(capi):
    a = capi.generate_mask_host(96, 80, seed=25, land_frac=0.45)
    b = capi.generate_mask_host(96, 80, seed=25, land_frac=0.45)
    assert np.array_equal(a, b)
    assert set(np.unique(a).tolist()) <= {0, 1}
    land = 1.0 - a.mean()
    assert 0.25 < land < 0.65, land
    # a row shard is the same rows of the global mask
    part = capi.generate_mask_host(96, 80, seed=25, land_frac=0.45, y_begin=30, y_count=20)
    assert np.array_equal(part, a[30:50])
    c = capi.generate_mask_host(96, 80, seed=26, land_frac=0.45)
    assert not np.array_equal(a, c)
    assert capi.generate_mask_host(16, 16, 1, 0.0).all()
    assert not capi.generate_mask_host(16, 16, 1, 1.0).any()
