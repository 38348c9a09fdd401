"""GPU parity at BASELINE.json's full sizes (configs C4 and C5).

The masks are synthesised on the device (bit-identical to the host generator, see
test_gpu_parity.test_device_generator_matches_host), decomposed through the C ABI, and checked
  * bit-exactly against the CPU oracle (both formulations at C4; the histogram formulation at C5,
    where the dot-based one would need tens of GB for Zoltan's dot arrays), and
  * through size-independent properties of a rectilinear decomposition: the boxes tile the
    domain as x-sorted strips of y-sorted parts, every ocean cell is labelled with the box that
    contains it and land with -1, the part loads are the label counts and sum to the ocean count,
    the neighbour relation is symmetric (p lists q on its LEFT with halo h <=> q lists p on its
    RIGHT with halo h; same for BOTTOM / TOP, interior and periodic), halo sizes are the overlap
    of the two boxes and halo starts index the neighbour's first shared cell (Partitioner.cpp:55-80).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LEFT, RIGHT, BOTTOM, TOP = range(4)


@pytest.fixture(scope="module")
def capi():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the -m gpu tests need one (the product has no CPU fallback)")
    from domain_decomp_b200 import capi
    capi.load()
    return capi


class _DevArray:
    """__cuda_array_interface__ view of an int32 device buffer owned by the library"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def decompose_on_device(capi, nx, ny, P, land, seed, px, py):
    import torch
    dev = torch.device("cuda", 0)
    d_mask = torch.empty((ny, nx), dtype=torch.int32, device=dev)
    h = capi.Handle(0)
    try:
        h.generate_mask_device(d_mask.data_ptr(), nx, ny, seed, land)
        h.set_mask_device(d_mask.data_ptr(), nx, ny)
        h.partition(P, bool(px), bool(py))
        out = {
            "boxes": h.boxes(),
            "stats": h.stats(),
            "loads": h.part_loads(),
            "counts": [[h.neighbour_counts(e, per) for e in range(4)] for per in range(2)],
            "nbr": [[h.neighbours(e, per) for e in range(4)] for per in range(2)],
        }
        # device-side facts about pid / mask before anything is copied to the host
        # the library's device-resident pid, viewed as a torch tensor and copied before the handle goes away
        d_pid = torch.as_tensor(_DevArray(h.pid_device(), (ny, nx)), device=dev).clone()
        torch.cuda.synchronize()
        out["d_mask"], out["d_pid"] = d_mask, d_pid
    finally:
        h.close()
    return out


def check_tiling(boxes, nx, ny):
    """the boxes are x-sorted full-height strips of y-sorted parts, in ascending part order"""
    x0, y0, ex, ey = (boxes[:, i].astype(np.int64) for i in range(4))
    assert (ex >= 0).all() and (ey >= 0).all()
    assert int((ex * ey).sum()) == nx * ny
    p, P, x_next = 0, len(boxes), 0
    strips = []
    while p < P:
        assert x0[p] == x_next, (p, x0[p], x_next)
        q, y_next = p, 0
        while q < P and x0[q] == x0[p] and ex[q] == ex[p] and y0[q] == y_next:
            y_next += ey[q]
            q += 1
            if y_next == ny:
                break
        assert y_next == ny, ("strip starting at part %d does not reach the top" % p, y_next)
        strips.append((p, q, int(x0[p]), int(x0[p] + ex[p])))
        x_next = x0[p] + ex[p]
        p = q
    assert x_next == nx
    return strips


def expected_pid_on_device(boxes, strips, d_mask):
    """labelling restated with torch: pid = ocean ? part whose box contains (x, y) : -1"""
    import torch
    ny, nx = d_mask.shape
    dev = d_mask.device
    S = len(strips)
    strip_of_col = torch.empty(nx, dtype=torch.int64, device=dev)
    rowpart = torch.empty((S, ny), dtype=torch.int32, device=dev)
    for s, (p, q, xa, xb) in enumerate(strips):
        strip_of_col[xa:xb] = s
        ys = torch.tensor(boxes[p:q, 1] + boxes[p:q, 3], dtype=torch.int64, device=dev)  # first row after each part
        rowpart[s] = (p + torch.searchsorted(ys, torch.arange(ny, device=dev), right=True)).to(torch.int32)
    out = torch.empty((ny, nx), dtype=torch.int32, device=dev)
    step = 2048  # rows per slab keeps the temporaries small
    for ya in range(0, ny, step):
        yb = min(ny, ya + step)
        lab = rowpart[:, ya:yb].t()[:, strip_of_col]  # [rows, nx]
        out[ya:yb] = torch.where(d_mask[ya:yb] > 0, lab, torch.full_like(lab, -1))
    return out


def check_neighbours(out, boxes, nx, ny, px, py):
    P = len(boxes)
    x1, y1 = boxes[:, 0].astype(np.int64), boxes[:, 1].astype(np.int64)
    x2, y2 = x1 + boxes[:, 2], y1 + boxes[:, 3]
    for per in range(2):
        rel = []
        for e in range(4):
            cnt = out["counts"][per][e].astype(np.int64)
            ids, halos, starts = (a.astype(np.int64) for a in out["nbr"][per][e])
            assert cnt.sum() == len(ids) == len(halos) == len(starts)
            me = np.repeat(np.arange(P, dtype=np.int64), cnt)
            # ids ascending inside every part's list (std::map order, Partitioner.cpp:98-110)
            same = me[1:] == me[:-1]
            assert (ids[1:][same] > ids[:-1][same]).all()
            # the literal edge test (Partitioner.cpp:20-53) and halo = overlap (DomainUtils.cpp:15-35)
            wx, wy = (nx if per and px else 0), (ny if per and py else 0)
            if e == LEFT:
                assert (x1[me] == x2[ids] - wx).all()
            elif e == RIGHT:
                assert (x2[me] == x1[ids] + wx).all()
            elif e == BOTTOM:
                assert (y1[me] == y2[ids] - wy).all()
            else:
                assert (y2[me] == y1[ids] + wy).all()
            w2 = x2[ids] - x1[ids]
            h2 = y2[ids] - y1[ids]
            if e in (LEFT, RIGHT):
                ov = np.minimum(y2[me], y2[ids]) - np.maximum(y1[me], y1[ids])
                dy = np.maximum(y1[me], y1[ids]) - y1[ids]
                st = (dy + 1) * w2 - 1 if e == LEFT else dy * w2
            else:
                ov = np.minimum(x2[me], x2[ids]) - np.maximum(x1[me], x1[ids])
                dx = np.maximum(x1[me], x1[ids]) - x1[ids]
                st = (h2 - 1) * w2 + dx if e == BOTTOM else dx
            assert (halos == ov).all() and (halos > 0).all()
            assert (starts == st).all()
            if not per:
                assert (ids != me).all()
            rel.append(set(zip(me.tolist(), ids.tolist(), halos.tolist())))
        # symmetry: LEFT of p holds (q, h)  <=>  RIGHT of q holds (p, h); BOTTOM / TOP alike
        assert rel[LEFT] == {(q, p, h) for (p, q, h) in rel[RIGHT]}
        assert rel[BOTTOM] == {(q, p, h) for (p, q, h) in rel[TOP]}
        if per:  # get_neighbour_info_periodic's filter (Partitioner.cpp:112-126)
            if not px:
                assert not rel[LEFT] and not rel[RIGHT]
            if not py:
                assert not rel[BOTTOM] and not rel[TOP]
    cut = sum(int(out["nbr"][0][e][1].astype(np.int64).sum()) for e in range(4))
    assert cut == out["stats"]["edge_cut"]


def check_properties(out, nx, ny, P, px, py):
    import torch
    boxes = out["boxes"]
    strips = check_tiling(boxes, nx, ny)
    assert out["stats"]["nstrips"] == len(strips)
    d_mask, d_pid = out["d_mask"], out["d_pid"]
    want = expected_pid_on_device(boxes, strips, d_mask)
    assert bool(torch.equal(want, d_pid)), "pid is not 'ocean ? containing box : -1'"
    del want
    n_ocean = int((d_mask > 0).sum().item())
    assert out["stats"]["n_ocean"] == n_ocean
    loads = torch.bincount(d_pid[d_pid >= 0].to(torch.int64), minlength=P).cpu().numpy()
    assert loads.tolist() == out["loads"].tolist()
    assert int(loads.sum()) == n_ocean
    assert out["stats"]["load_max"] == int(loads.max()) and out["stats"]["load_min"] == int(loads.min())
    assert out["stats"]["changes"] == 1
    # RCB balances the parts: with unit weights and thousands of cells per part the heaviest part
    # stays within a few columns / rows of the mean
    assert loads.max() <= 1.10 * n_ocean / P
    check_neighbours(out, boxes, nx, ny, px, py)


def assert_same_as_oracle(out, pid_host, o):
    assert out["boxes"].tolist() == o.boxes.tolist()
    assert np.array_equal(pid_host, o.pid)
    assert out["stats"]["changes"] == o.changes
    assert out["stats"]["median_iters"] == o.median_iters
    for per in range(2):
        for e in range(4):
            assert np.array_equal(out["counts"][per][e], o.nbr.counts[per][e]), (per, e)
            for got, want in zip(out["nbr"][per][e], (o.nbr.ids[per][e], o.nbr.halos[per][e], o.nbr.starts[per][e])):
                assert np.array_equal(got, want), (per, e)


def test_c4_arctic1km_4096_parts(capi, oracle):
    """BASELINE config 4: 8192 x 8192, ~60 % land, 4096 parts -- both oracle formulations"""
    nx, ny, P, land, seed = 8192, 8192, 4096, 0.60, 1
    out = decompose_on_device(capi, nx, ny, P, land, seed, 0, 0)
    check_properties(out, nx, ny, P, 0, 0)
    mask = out["d_mask"].cpu().numpy()
    pid = out["d_pid"].cpu().numpy()
    assert 0.3 < (mask > 0).mean() < 0.5
    for use_hist in (True, False):
        o = oracle.partition(mask, P, False, False, use_hist=use_hist)
        assert_same_as_oracle(out, pid, o)


@pytest.mark.parametrize("P", [3000, 4097])
def test_c4_part_counts_that_are_not_powers_of_two(capi, oracle, P):
    """the same 8192 x 8192 mask into 3000 and 4097 parts: uneven part counts at every level of the RCB tree
    (ceil(n/2) | floor(n/2) splits, leaves at different depths, targets that are not W/2), both periodic"""
    nx, ny, land, seed = 8192, 8192, 0.60, 1
    out = decompose_on_device(capi, nx, ny, P, land, seed, 1, 1)
    check_properties(out, nx, ny, P, 1, 1)
    mask = out["d_mask"].cpu().numpy()
    pid = out["d_pid"].cpu().numpy()
    assert_same_as_oracle(out, pid, oracle.partition(mask, P, True, True, use_hist=True))


def test_c5_1km_global_16384_parts(capi, oracle):
    """BASELINE config 5: 32768 x 32768 (1.07 G cells), 16384 parts, periodic in x"""
    import psutil
    import torch
    if psutil.virtual_memory().available < 48 * 2**30:
        pytest.skip("needs ~16 GiB of host memory for the oracle's copy of mask and pid")
    if torch.cuda.mem_get_info(0)[0] < 24 * 2**30:
        pytest.skip("needs ~20 GiB of free device memory")
    nx, ny, P, land, seed = 32768, 32768, 16384, 0.45, 32
    out = decompose_on_device(capi, nx, ny, P, land, seed, 1, 0)
    check_properties(out, nx, ny, P, 1, 0)
    mask = out["d_mask"].cpu().numpy()
    pid = out["d_pid"].cpu().numpy()
    del out["d_mask"], out["d_pid"]
    torch.cuda.empty_cache()
    o = oracle.partition(mask, P, True, False, use_hist=True)
    assert_same_as_oracle(out, pid, o)
