#!/usr/bin/env python
"""Small decompositions through the C ABI, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
    compute-sanitizer --tool initcheck python scripts/sanitize_small.py

Covers every kernel of the path (vector and scalar variants, ragged widths, empty masks, P == 1,
P > columns, periodic combinations, the naive-block fallback, caller-supplied boxes) and checks each
result against the CPU oracle, so that a sanitizer-clean run is also a parity run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from domain_decomp_b200 import capi  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rng = np.random.default_rng(11)
    cases = []
    for (nx, ny, P, px, py) in [(6, 4, 3, 0, 0), (6, 4, 3, 1, 1), (30, 30, 4, 0, 0), (33, 17, 5, 1, 0), (128, 64, 16, 0, 1),
                                (257, 130, 7, 1, 1), (64, 64, 1, 1, 1), (5, 40, 12, 0, 0), (96, 64, 5, 0, 0)]:
        m = (rng.random((ny, nx)) < 0.6).astype(np.int32)
        cases.append((m, P, px, py))
    cases.append((np.zeros((64, 96), dtype=np.int32), 5, 0, 0))  # all land -> naive blocks
    cases.append((np.ones((60, 60), dtype=np.int32), 6, 1, 1))  # nothing moves
    cases.append((capi.generate_mask_host(528, 522, 25, 0.45), 64, 1, 1))
    cases.append((capi.generate_mask_host(1024, 300, 3, 0.5), 96, 0, 0))
    bad = 0
    with capi.Handle(0) as h:
        for i, (m, P, px, py) in enumerate(cases):
            h.set_mask_host(np.ascontiguousarray(m))
            h.partition(P, bool(px), bool(py))
            o = orc.partition(m, P, bool(px), bool(py), use_hist=True)
            ok = h.boxes().tolist() == o.boxes.tolist() and np.array_equal(h.pid_host(), o.pid)
            for per in range(2):
                for e in range(4):
                    a, b, c = h.neighbours(e, per)
                    ok &= a.tolist() == o.nbr.ids[per][e].tolist() and b.tolist() == o.nbr.halos[per][e].tolist()
                    ok &= c.tolist() == o.nbr.starts[per][e].tolist()
            print("case %d %dx%d P=%d px=%d py=%d: %s" % (i, m.shape[1], m.shape[0], P, px, py, "ok" if ok else "MISMATCH"))
            bad += 0 if ok else 1
        boxes = np.array([[0, 0, 10, 10], [10, 0, 10, 5], [10, 5, 10, 5], [0, 10, 20, 10]], dtype=np.int32)
        h.neighbours_from_boxes(boxes, 20, 20, True, True)
        want = orc.neighbours(boxes, 20, 20, True, True)
        for per in range(2):
            for e in range(4):
                a, b, c = h.neighbours(e, per)
                if a.tolist() != want.ids[per][e].tolist() or b.tolist() != want.halos[per][e].tolist():
                    bad += 1
    print("SANITIZE RUN", "OK" if bad == 0 else "FAILED (%d)" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
