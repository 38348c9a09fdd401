"""Result checks that do not need the oracle: a digest of the replicated tables of a decomposition and a
device-side restatement of the labelling rule.  bench.py uses them (outside its timed regions) to assert
that every rank of a multi-GPU run ended with the same boxes / neighbour tables as the one-GPU run, and that
its pid rows are the labelling of exactly those boxes -- the GPU counterpart of the four MPI_Allgather calls
every rank of the reference has to agree on (Partitioner.cpp:378-388).
"""
from __future__ import annotations

import hashlib

import numpy as np


def result_digest(boxes, counts, nbr, loads, changes) -> str:
    """sha256 over the part boxes [P, 4] (x0, y0, extent x, extent y), the eight neighbour tables
    (list = periodic * 4 + edge: counts[P], ids, halo sizes, halo starts), the part loads and `changes`"""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(boxes, dtype=np.int32).tobytes())
    for per in range(2):
        for e in range(4):
            h.update(np.ascontiguousarray(counts[per][e], dtype=np.int32).tobytes())
            for a in nbr[per][e]:
                h.update(np.ascontiguousarray(a, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(loads, dtype=np.int64).tobytes())
    h.update(np.int32(changes).tobytes())
    return h.hexdigest()


def handle_digest(h) -> str:
    """the digest of what a capi.Handle holds after ddc_partition (all getters of the C ABI)"""
    counts = [[h.neighbour_counts(e, per) for e in range(4)] for per in range(2)]
    nbr = [[h.neighbours(e, per) for e in range(4)] for per in range(2)]
    return result_digest(h.boxes(), counts, nbr, h.part_loads(), h.stats()["changes"])


def strips_of(boxes: np.ndarray, nx: int, ny: int):
    """the boxes as x-sorted full-height strips of y-sorted parts: [(first part, end part, x begin, x end)];
    raises AssertionError when they do not tile the domain that way"""
    x0, y0, ex, ey = (boxes[:, i].astype(np.int64) for i in range(4))
    assert (ex >= 0).all() and (ey >= 0).all() and int((ex * ey).sum()) == nx * ny, "boxes do not cover the domain"
    p, P, x_next, strips = 0, len(boxes), 0, []
    while p < P:
        assert x0[p] == x_next, "strip of part %d does not start where the previous one ended" % p
        q, y_next = p, 0
        while q < P and x0[q] == x0[p] and ex[q] == ex[p] and y0[q] == y_next:
            y_next += ey[q]
            q += 1
            if y_next == ny:
                break
        assert y_next == ny, "strip starting at part %d does not reach the top" % p
        strips.append((p, q, int(x0[p]), int(x0[p] + ex[p])))
        x_next = x0[p] + ex[p]
        p = q
    assert x_next == nx
    return strips


def pid_rows_match_boxes(boxes: np.ndarray, nx: int, ny: int, d_mask, d_pid, y_begin: int):
    """pid = ocean ? part whose box contains (x, y) : -1 (ZoltanPartitioner.cpp:201-219, by box lookup), checked
    with torch on the device for the rows [y_begin, y_begin + rows) a rank holds.  changes == 0 decompositions
    (naive blocks, which need not be strips) are checked box by box instead.  Returns (ok, per-part label counts)."""
    import torch
    rows = d_mask.shape[0]
    dev = d_mask.device
    P = len(boxes)
    counts = torch.bincount((d_pid.reshape(-1)[d_pid.reshape(-1) >= 0]).to(torch.int64), minlength=P)
    try:
        strips = strips_of(boxes, nx, ny)
    except AssertionError:
        strips = None
    ok = True
    if strips is not None:
        S = len(strips)
        strip_of_col = torch.empty(nx, dtype=torch.int64, device=dev)
        rowpart = torch.empty((S, rows), dtype=torch.int32, device=dev)
        ys_local = torch.arange(y_begin, y_begin + rows, device=dev)
        for s, (p, q, xa, xb) in enumerate(strips):
            strip_of_col[xa:xb] = s
            ends = torch.tensor(boxes[p:q, 1] + boxes[p:q, 3], dtype=torch.int64, device=dev)
            rowpart[s] = (p + torch.searchsorted(ends, ys_local, right=True)).to(torch.int32)
        step = 1024
        for ya in range(0, rows, step):
            yb = min(rows, ya + step)
            lab = rowpart[:, ya:yb].t()[:, strip_of_col]
            want = torch.where(d_mask[ya:yb] > 0, lab, torch.full_like(lab, -1))
            if not torch.equal(want, d_pid[ya:yb]):
                ok = False
                break
    else:
        want = torch.full_like(d_pid, -1)
        for p in range(P):
            xa, ya, ex, ey = (int(v) for v in boxes[p])
            a, b = max(ya, y_begin) - y_begin, min(ya + ey, y_begin + rows) - y_begin
            if ex > 0 and b > a:
                want[a:b, xa:xa + ex] = p
        want = torch.where(d_mask > 0, want, torch.full_like(want, -1))
        ok = bool(torch.equal(want, d_pid))
    return ok, counts
