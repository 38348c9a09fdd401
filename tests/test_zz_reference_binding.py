"""The drop-in on a GPU: the reference's own Grid.cpp / Partitioner.cpp / DomainUtils.cpp (compiled where they
lie into oracle/_ref/libref_binding.so on the build machine) + integration/reference_binding (the
CudaRcbPartitioner a maintainer adds to the reference tree) + libddc_cuda.so.  P thread-ranks run the
reference's code path `Grid::create -> partitioner->partition -> save_mask / save_metadata` exactly as
main.cpp:78-94 does, with the Zoltan partitioner swapped for the CUDA one, and must write the reference's
golden files.  (Runs last: it needs the prebuilt oracle/_ref, which only exists where the reference checkout
was available at build time.)"""
import numpy as np
import pytest

from test_reference_binding import check_against_oracle, check_goldens

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(oracle):
    from domain_decomp_b200 import capi
    try:
        capi.Handle(0).close()
    except Exception as e:  # no torch needed here: the C ABI itself says whether there is a device
        pytest.fail("no CUDA device: the -m gpu tests need one (the product has no CPU fallback): %s" % e)
    if oracle.ref_binding_lib(cpu=False) is None:
        pytest.skip("oracle/_ref/libref_binding.so not built (no reference checkout on the build machine)")
    return oracle


@pytest.mark.parametrize("case", ["test_1", "test_2", "test_1_px", "test_1_py", "test_1_px_py"])
def test_reference_code_with_cuda_partitioner_writes_the_goldens(goldens, ref, case):
    check_goldens(goldens, lambda *a, **k: ref.ref_binding_run(*a, cpu=False, **k), case)


def test_reference_code_with_cuda_partitioner_random(ref):
    from domain_decomp_b200 import capi
    run = lambda *a, **k: ref.ref_binding_run(*a, cpu=False, **k)
    rng = np.random.default_rng(43)
    for (nx, ny, P) in [(30, 30, 4), (37, 21, 6), (64, 48, 8), (96, 64, 12)]:
        mask = (rng.random((ny, nx)) >= 0.4).astype(np.int32)
        check_against_oracle(ref, run, mask, P, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
    check_against_oracle(ref, run, np.zeros((4, 6), dtype=np.int32), 2, False, False)
    check_against_oracle(ref, run, capi.generate_mask_host(528, 522, 25, 0.45), 16, True, True)
