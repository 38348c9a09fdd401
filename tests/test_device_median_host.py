"""The scalar core of the CUDA cut kernels, fuzzed on the CPU.

domain_decomp_b200/csrc/ddc_median.cuh (median_boundary, the three-level bit map queries, rcb_walk) contains
no thread / warp / block primitive, so the very same source compiles as host C++ (oracle/emu_median_harness.cpp,
-DDDC_HOST_EMU).  The device code is not a transliteration of Zoltan's loop: weights are integers, the tests
against the targets are integer tests, the interpolated guess is formed in FP32 with a guard band and only
falls back to the exact FP64 sequence near integers.  Here every median -- boundary AND iteration count --
is compared with the oracle's literal double-precision loop (ddc_oracle.c: find_median_hist) on millions of
random histograms (dense, sparse, ties, single dots, coastline-like; up to 6 M bins so that the FP64 fallback
for ranges beyond 2^22 runs too), with and without the bit map, and with the fast division perturbed by
+-2 ulp, the documented accuracy of __fdividef; the barrier-free tree walk is compared with the oracle's
level-by-level recursion."""
import ctypes as C

import pytest

SHAPES = {0: "dense", 1: "sparse", 2: "ties", 3: "single dots", 4: "coastline"}


@pytest.fixture(scope="module")
def emu(oracle):
    return oracle.median_emu_lib()


def fuzz(emu, seed, histograms, queries, nmax, shape, ulps, bitmap):
    bad = (C.c_longlong * 10)()
    n = C.c_longlong()
    m = emu.emu_fuzz(seed, histograms, queries, nmax, shape, ulps, int(bitmap), bad, C.byref(n))
    return m, n.value, list(bad)


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("ulps", [0, 2, -2])
def test_device_median_equals_oracle_small_histograms(emu, shape, ulps):
    for bitmap in (True, False):
        m, n, bad = fuzz(emu, 1000 + shape, 3000, 50, 600, shape, ulps, bitmap)
        assert n > 150000
        assert m == 0, "%s, ulps %+d, bit map %s: first mismatch {kind,n,c0,c1,nlo,parts,got,want,it,want_it} = %s" % (
            SHAPES[shape], ulps, bitmap, bad)


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_device_median_equals_oracle_multi_tile_histograms(emu, shape):
    """up to 70000 bins: three tiles of the bit map, the sizes of the wide-grid configs"""
    for ulps in (0, 2, -2):
        m, n, bad = fuzz(emu, 2000 + shape, 300, 80, 70000, shape, ulps, True)
        assert m == 0, (SHAPES[shape], ulps, bad)


@pytest.mark.parametrize("shape", [1, 4])
def test_device_median_equals_oracle_beyond_fp32_ranges(emu, shape):
    """ranges of 2^22 bins and more take the exact FP64 guess"""
    for ulps in (0, 2):
        m, n, bad = fuzz(emu, 3000 + shape, 6, 200, 6000000, shape, ulps, True)
        assert m == 0, (SHAPES[shape], ulps, bad)


def test_the_fuzz_sees_a_broken_guard(emu):
    """a division that is wrong by far more than the guard band allows must show up as mismatches"""
    m, n, bad = fuzz(emu, 5, 2000, 40, 3000, 4, 50000, True)
    assert m > 0


@pytest.mark.parametrize("degenerate", [0, 1])
@pytest.mark.parametrize("maxn,maxstrips,maxparts,cases", [(5, 6, 6, 20000), (12, 4, 4, 20000), (40, 8, 8, 10000),
                                                             (300, 16, 24, 1500)])
def test_device_neighbour_search_equals_oracle(emu, degenerate, maxn, maxstrips, maxparts, cases):
    """K7's structured search (csrc/ddc_neighbours.cuh: one thread per (list, part), binary searches over the
    x-sorted strips and the y-sorted parts) as host code against the oracle's literal O(P^2) discovery: counts,
    ids, halo sizes, halo starts of all eight lists and the edge cut, on random tilings -- degenerate = 1 lets
    boundaries coincide (zero-width strips, zero-height parts: more parts than non-empty columns / rows)"""
    bad = (C.c_longlong * 10)()
    n = C.c_longlong()
    m = emu.emu_fuzz_neighbours(11 + degenerate, cases, maxn, maxstrips, maxparts, degenerate, bad, C.byref(n))
    assert n.value == 8 * cases
    assert m == 0, "first mismatch {NX, NY, P, list, part, what} = %s" % list(bad)[:6]
