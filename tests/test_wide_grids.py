"""Grids whose histograms do not fit the cut kernels' shared memory, or whose rows hold 65536 cells or
more: the global-memory prefix path of K2 / K4 (no bit map of non-empty bins: the nearest-non-empty-bin
queries fall back to binary searches over the prefix sums), several prefix tiles per histogram, and
32-bit strip row counts.  Bit-exact against the histogram oracle like every other parity test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the -m gpu tests need one (the product has no CPU fallback)")
    from domain_decomp_b200 import capi
    capi.load()
    return capi


def run_and_compare(capi, oracle, nx, ny, P, px, py, seed, land):
    import torch
    h = capi.Handle(0)
    try:
        # the mask is synthesised on the device (bit-identical to the host generator) and copied back for the oracle
        d_mask = torch.empty((ny, nx), dtype=torch.int32, device="cuda:0")
        h.generate_mask_device(d_mask.data_ptr(), nx, ny, seed, land)
        h.set_mask_device(d_mask.data_ptr(), nx, ny)
        for _ in range(2):  # the second call runs with the plan of the first
            h.partition(P, px, py)
        mask = d_mask.cpu().numpy()
        assert 0.05 < (mask > 0).mean() < 0.95
        o = oracle.partition(mask, P, px, py, use_hist=True)
        assert h.boxes().tolist() == o.boxes.tolist()
        assert np.array_equal(h.pid_host(), o.pid)
        st = h.stats()
        assert st["changes"] == o.changes and st["median_iters"] == o.median_iters
        assert h.part_loads().tolist() == oracle.part_loads(o.pid, P).tolist()
        for per in range(2):
            for e in range(4):
                assert h.neighbour_counts(e, per).tolist() == o.nbr.counts[per][e].tolist()
                ids, halos, starts = h.neighbours(e, per)
                assert ids.tolist() == o.nbr.ids[per][e].tolist()
                assert halos.tolist() == o.nbr.halos[per][e].tolist()
                assert starts.tolist() == o.nbr.starts[per][e].tolist()
    finally:
        h.close()


@pytest.mark.parametrize("nx,ny,P", [(70000, 24, 12), (66001, 31, 7), (131072, 16, 64), (66000, 4100, 64)])
def test_rows_of_65536_cells_or_more(capi, oracle, nx, ny, P):
    """NX >= 65536: the column histogram (NX + 1 prefix sums + bit map) exceeds 227 KB of shared memory, so K2
    scans into global memory, 3 to 4 tiles of 32768 bins; the last case (270 M cells, 5 x levels then one
    y level) also goes through the 32-bit strip row counts"""
    run_and_compare(capi, oracle, nx, ny, P, True, False, nx % 97, 0.4)


@pytest.mark.parametrize("nx,ny,P", [(40, 60000, 24), (24, 70001, 9), (300, 57000, 96), (2000, 57000, 64)])
def test_columns_taller_than_shared_memory(capi, oracle, nx, ny, P):
    """NY above ~55000: the row histogram of a strip does not fit shared memory, K4 scans into global memory"""
    run_and_compare(capi, oracle, nx, ny, P, False, True, ny % 89, 0.5)


def test_histogram_of_exactly_one_tile_and_one_more(capi, oracle):
    """32768 bins is one prefix tile, 32769 two; 49152 still fits shared memory with two tiles"""
    for nx, ny, P in [(32768, 20, 16), (32769, 20, 16), (49152, 12, 8), (16, 32769, 8)]:
        run_and_compare(capi, oracle, nx, ny, P, False, False, 3, 0.3)


def test_around_the_shared_memory_limit_of_the_cut_kernels(capi, oracle):
    """The cut kernels keep the prefix sums, the bit map and the sets of two RCB levels in dynamic shared memory and a
    little more statically; the library asks the kernel for its static size (cudaFuncGetAttributes) before it chooses
    the shared-memory variant.  Extents just below and above the limit -- in x and in y -- must both decompose
    (round 1 reserved a fixed 1024 bytes and failed to launch in a window of a few columns)."""
    lim, static = 232448, 1056 + 64

    def need(n):
        tiles = (n + 32767) // 32768
        return 4 * ((((n + 1) + 3) & ~3) + tiles * (1024 + 32 + 1)) + (2 * 1024 * 16 + 16)

    n = 40000
    while need(n) + static <= lim:
        n += 1
    for nx in (n - 9, n - 3, n - 1, n, n + 2, n + 8):
        run_and_compare(capi, oracle, nx, 12, 8, False, False, 5, 0.3)
    for ny in (n - 2, n + 1):
        run_and_compare(capi, oracle, 16, ny, 8, False, False, 7, 0.3)
