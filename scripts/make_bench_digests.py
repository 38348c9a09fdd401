#!/usr/bin/env python
"""Golden digests of bench.py's workloads from the CPU oracle (TEST INFRASTRUCTURE).

    python scripts/make_bench_digests.py [workload ...]      -> tests/golden/bench_digests.json

For every workload: the synthetic mask from the oracle's restatement of the generator, the whole path through
oracle/ddc_oracle.c (histogram formulation; up to 8192 x 8192 also the literal dot-based Zoltan loop, which
must give the same tables), and the sha256 that domain_decomp_b200/verify.result_digest computes over the
part boxes, the eight neighbour tables, the part loads and `changes`.  bench.py compares the digest of what
it timed on the GPU(s) with these; tests/test_gpu_parity.py does the same for the small workloads.
C5 (32768 x 32768) needs about 10 GB of host memory and a few minutes.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import GOLDEN_DIGESTS, WORKLOADS  # noqa: E402
from domain_decomp_b200 import verify  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def digest_of(name):
    nx, ny, P, land, seed, px, py = WORKLOADS[name]
    t = time.time()
    mask = orc.generate_mask(nx, ny, seed, land)
    o = orc.partition(mask, P, bool(px), bool(py), use_hist=True)
    loads = orc.part_loads(o.pid, P)
    d = verify.result_digest(o.boxes, o.nbr.counts, [[(o.nbr.ids[per][e], o.nbr.halos[per][e], o.nbr.starts[per][e])
                                                      for e in range(4)] for per in range(2)], loads, o.changes)
    if nx * ny <= 8192 * 8192:
        orc.set_threads(orc.max_threads())
        o2 = orc.partition(mask, P, bool(px), bool(py), use_hist=False)
        assert np.array_equal(o.boxes, o2.boxes) and np.array_equal(o.pid, o2.pid), "dot and histogram formulations differ"
    return {"digest": d, "nx": nx, "ny": ny, "parts": P, "land_frac": land, "seed": seed, "periodic_x": px, "periodic_y": py,
            "n_ocean": int((mask > 0).sum()), "changes": int(o.changes), "median_iters": int(o.median_iters),
            "seconds": round(time.time() - t, 1)}


def main():
    names = sys.argv[1:] or [n for n in WORKLOADS]
    try:
        with open(GOLDEN_DIGESTS) as f:
            out = json.load(f)
    except Exception:
        out = {}
    for n in names:
        out[n] = digest_of(n)
        print(n, out[n], flush=True)
        with open(GOLDEN_DIGESTS, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
