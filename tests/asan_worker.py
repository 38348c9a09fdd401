"""Runs a handful of decompositions through the AddressSanitizer build of the kernel emulation
(oracle/libddc_emu_asan.so); started by test_kernels_asan.py with libasan preloaded."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.oracle as orc  # noqa: E402

L = C.CDLL(os.path.join(ROOT, "oracle", "libddc_emu_asan.so"))
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
L.emu_partition.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                            i32p, i32p, i32p, i32p, C.c_long, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")]
L.emu_partition.restype = C.c_int
L.emu_last_error.restype = C.c_char_p
orc._emu = L  # orc.emu_partition() now drives the sanitizer build

from domain_decomp_b200 import capi  # noqa: E402

rng = np.random.default_rng(3)
cases = [(capi.generate_mask_host(100, 40, 7, 0.5), 12, True, False, dict(ranks=2)),
         (capi.generate_mask_host(67, 31, 2, 0.4), 7, False, True, dict(ranks=1)),  # ragged width: scalar paths
         (capi.generate_mask_host(150, 40, 11, 0.45), 12, True, False, dict(ranks=2, strip_k=4)),
         (capi.generate_mask_host(60, 50, 5, 0.5), 8, False, True, dict(ranks=2, smem_limit=2048)),
         (np.ones((24, 24), dtype=np.int32), 4, True, True, dict(ranks=1)),  # nothing moved: K5 rebuilds the tables
         ((rng.random((5, 3)) < 0.5).astype(np.int32), 9, True, True, dict(ranks=1))]  # more parts than columns
for mask, P, px, py, kw in cases:
    d, _ = orc.emu_partition(mask, P, px, py, **kw)
    o = orc.partition(mask, P, px, py, use_hist=True)
    assert d.boxes.tolist() == o.boxes.tolist() and np.array_equal(d.pid, o.pid), (mask.shape, P, kw)
print("ASAN RUN OK: %d decompositions" % len(cases))
