"""The drop-in, end to end, on the CPU: the reference's own Grid.cpp / Partitioner.cpp / DomainUtils.cpp (compiled
where they lie, oracle/_ref) + integration/reference_binding/CudaRcbPartitioner.cpp -- the one file a maintainer
adds to the reference tree, written against the reference's headers and include/ddc.h -- with, behind the C ABI,
  * backend True:  the CPU oracle answering the binding's ddc_* calls (oracle/ddc_oracle_stub.c), or
  * backend "emu": the PRODUCT's own C-ABI implementation and kernels on the host emulation of CUDA
                   (oracle/libddc_cuda_emu.so) -- the whole drop-in stack, reference code on top, without a GPU.
tests/test_zz_reference_binding.py runs the same thing on a GPU with the real CUDA library."""
import numpy as np
import pytest

from conftest import golden_mask
from test_reference_hostpath import EDGES, per_part, sane_blocks


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.ref_binding_lib(cpu=True) is None or oracle.ref_binding_lib(cpu="emu") is None:
        pytest.skip("oracle/_ref/libref_binding_{cpu,emu}.so not built (no reference checkout here)")
    return oracle


def check_goldens(goldens, run, case):
    G = goldens["integration"][case]
    inp = goldens["inputs"][G["input"]]
    r = run(golden_mask(goldens, G["input"]), G["P"], bool(G["px"]), bool(G["py"]), xdim=inp["xdim"],
            ydim=inp["ydim"], maskname=inp["mask_name"])
    meta, mfile = r["files"]["metadata"], r["files"]["mask"]
    assert dict(meta["dims"]) == G["dims"]
    got = {name: vals for (grp, name), (dims, vals) in meta["vars"].items()}
    assert {k: v for k, v in got.items() if v} == G["metadata"]
    assert mfile["vars"][("/", "pid")] == ("(y,x)", list(G["pid"])) and mfile["atts"] == {"num_processes": G["P"]}
    assert not meta["unwritten"] and not mfile["unwritten"]


def check_against_oracle(orc, run, mask, P, px, py):
    ny, nx = mask.shape
    o = orc.partition(mask, P, px, py, use_hist=True)
    r = run(mask, P, px, py)
    ctx = (nx, ny, P, px, py)
    for p in range(P):
        assert r["ranks"][p]["box"] == o.boxes[p].tolist(), ctx
        for per in range(2):
            for e in range(4):
                want = per_part(o.nbr, P, per, e)[p] if P > 1 else []
                assert r["ranks"][p]["nbr"][per][e] == want, (ctx, p, per, e)
    meta = {name: vals for (grp, name), (dims, vals) in r["files"]["metadata"]["vars"].items()}
    for i, key in enumerate(("domain_x", "domain_y", "domain_extent_x", "domain_extent_y")):
        assert meta[key] == o.boxes[:, i].tolist(), ctx
    for per, sfx in ((0, ""), (1, "_periodic")):
        for e, name in enumerate(EDGES):
            if P == 1:
                continue
            assert meta[name + "_neighbours" + sfx] == o.nbr.counts[per][e].tolist(), (ctx, name, sfx)
            assert meta[name + "_neighbour_ids" + sfx] == o.nbr.ids[per][e].tolist(), (ctx, name, sfx)
            assert meta[name + "_neighbour_halos" + sfx] == o.nbr.halos[per][e].tolist(), (ctx, name, sfx)
            assert meta[name + "_neighbour_halo_starts" + sfx] == o.nbr.starts[per][e].tolist(), (ctx, name, sfx)
    assert r["files"]["mask"]["vars"][("/", "pid")][1] == o.pid.ravel().tolist(), ctx
    assert not r["files"]["metadata"]["unwritten"] and not r["files"]["mask"]["unwritten"]


@pytest.mark.parametrize("backend", [True, "emu"])
@pytest.mark.parametrize("case", ["test_1", "test_2", "test_1_px", "test_1_py", "test_1_px_py"])
def test_binding_reproduces_the_goldens(goldens, ref, case, backend):
    check_goldens(goldens, lambda *a, **k: ref.ref_binding_run(*a, cpu=backend, **k), case)


@pytest.mark.parametrize("backend,cases", [(True, 60), ("emu", 6)])
def test_binding_random_masks(ref, backend, cases):
    rng = np.random.default_rng(41)
    run = lambda *a, **k: ref.ref_binding_run(*a, cpu=backend, **k)
    done = 0
    while done < cases:
        nx, ny, P = int(rng.integers(2, 40)), int(rng.integers(2, 40)), int(rng.integers(1, 17 if backend is True else 7))
        if not sane_blocks(ref, P, nx, ny):
            continue
        mask = (rng.random((ny, nx)) >= rng.random() * 0.8).astype(np.int32) * int(rng.integers(1, 3))
        check_against_oracle(ref, run, mask, P, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        done += 1
    check_against_oracle(ref, run, np.zeros((4, 6), dtype=np.int32), 2, False, False)  # all land: naive blocks
    check_against_oracle(ref, run, np.ones((60, 60), dtype=np.int32), 6, True, True)  # nothing moves


def test_binding_ignore_mask(ref):
    """--ignore-mask: Grid reads no mask, the binding treats every cell as ocean"""
    mask = np.zeros((8, 12), dtype=np.int32)
    r = ref.ref_binding_run(mask, 4, cpu=True, ignore_mask=True)
    o = ref.partition(np.ones_like(mask), 4, use_hist=True)
    assert [r["ranks"][p]["box"] for p in range(4)] == o.boxes.tolist()
    assert r["files"]["mask"]["vars"][("/", "pid")][1] == o.pid.ravel().tolist()
