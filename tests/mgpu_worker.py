"""Multi-GPU parity worker: run under torchrun, one rank per GPU (NCCL), or with --emulate on
CPU/gloo to exercise only the host-side sharding logic.

    torchrun --nproc-per-node 2 tests/mgpu_worker.py

Every rank row-shards the same synthetic masks, calls the C ABI with its shard, and compares the
(replicated) boxes / neighbour tables and its own pid rows with the CPU oracle.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from domain_decomp_b200 import capi  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [  # nx, ny, P, land, seed, px, py
    (528, 522, 64, 0.45, 25, 1, 1),
    (1024, 777, 96, 0.5, 7, 0, 1),
    (2048, 2048, 256, 0.45, 3, 1, 0),
    (640, 5, 8, 0.3, 2, 0, 0),  # fewer rows than some shard counts want: empty shards
    (96, 64, 5, 1.0, 1, 0, 0),  # all land -> naive blocks
    (4096, 4096, 1024, 0.45, 3, 0, 0),
]


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    h = capi.Handle(local, rank, world, ids[0])
    failures = 0
    for mode in ("nccl", "peer"):
        if mode == "peer":
            # hand every rank's exchange buffer to every other rank (CUDA IPC handles through the host)
            mine = h.peer_export(4096, 4096, 1024)
            handles = [None] * world
            dist.all_gather_object(handles, mine)
            h.peer_import(handles)
        for (nx, ny, P, land, seed, px, py) in CASES:
            mask = capi.generate_mask_host(nx, ny, seed, land)
            yb, yc = capi.shard_rows(ny, world, rank)
            shard = np.ascontiguousarray(mask[yb:yb + yc])
            if yc == 0:
                shard = np.zeros((0, nx), dtype=np.int32)
            h.set_mask_host(shard, ny=ny, y_begin=yb)
            for rep in range(2):  # twice: both parities of the double-buffered exchange
                h.partition(P, bool(px), bool(py))
            o = orc.partition(mask, P, bool(px), bool(py), use_hist=True)
            ok = h.boxes().tolist() == o.boxes.tolist()
            ok &= np.array_equal(h.pid_host(), o.pid[yb:yb + yc])
            st = h.stats()
            ok &= st["exchange"] == (1 if mode == "nccl" else 2)
            ok &= st["changes"] == o.changes and st["n_ocean"] == int((mask > 0).sum())
            ok &= st["median_iters"] == o.median_iters
            for per in range(2):
                for e in range(4):
                    ok &= h.neighbour_counts(e, per).tolist() == o.nbr.counts[per][e].tolist()
                    a, b, c = h.neighbours(e, per)
                    ok &= a.tolist() == o.nbr.ids[per][e].tolist() and b.tolist() == o.nbr.halos[per][e].tolist()
                    ok &= c.tolist() == o.nbr.starts[per][e].tolist()
            ok &= h.part_loads().tolist() == orc.part_loads(o.pid, P).tolist()
            print("rank %d/%d %s case %dx%d P=%d: %s" % (rank, world, mode, nx, ny, P, "ok" if ok else "MISMATCH"),
                  flush=True)
            failures += 0 if ok else 1
    # the BASELINE sizes: the mask is synthesised on the device, the replicated tables are compared with the CPU
    # oracle's digest (tests/golden/bench_digests.json), the pid rows with the labelling rule restated in torch
    if "--big" in sys.argv:
        import json
        from domain_decomp_b200 import verify
        with open(os.path.join(ROOT, "tests", "golden", "bench_digests.json")) as f:
            golden = json.load(f)
        dev = torch.device("cuda", local)
        for name in ("C4_8192x8192_p4096", "C5_32768x32768_p16384"):
            g = golden[name]
            nx, ny, P, px, py = g["nx"], g["ny"], g["parts"], g["periodic_x"], g["periodic_y"]
            yb, yc = capi.shard_rows(ny, world, rank)
            d_mask = torch.empty((max(yc, 1), nx), dtype=torch.int32, device=dev)
            # a handle of its own, with exchange buffers exported for this size
            hh = capi.Handle(local, rank, world, None)  # peer-memory exchange only: no NCCL communicator
            handles = [None] * world
            dist.all_gather_object(handles, hh.peer_export(nx, ny, P))
            hh.peer_import(handles)
            hh.generate_mask_device(d_mask.data_ptr(), nx, ny, g["seed"], g["land_frac"], yb, yc)
            hh.set_mask_device(d_mask.data_ptr(), nx, ny, yb, yc)
            for rep in range(3):
                hh.partition(P, bool(px), bool(py))
            st = hh.stats()
            ok = st["exchange"] == 2 and verify.handle_digest(hh) == g["digest"] and st["median_iters"] == g["median_iters"]

            class _Dev:
                def __init__(self, ptr, shape):
                    self.__cuda_array_interface__ = {"shape": shape, "typestr": "<i4", "data": (int(ptr), False),
                                                     "version": 2, "strides": None}
            d_pid = torch.as_tensor(_Dev(hh.pid_device(), (max(yc, 1), nx)), device=dev)[:yc]
            pid_ok, counts = verify.pid_rows_match_boxes(hh.boxes(), nx, ny, d_mask[:yc], d_pid, yb)
            dist.all_reduce(counts)
            ok &= bool(pid_ok) and counts.cpu().tolist() == hh.part_loads().tolist()
            print("rank %d/%d peer case %s: %s" % (rank, world, name, "ok" if ok else "MISMATCH"), flush=True)
            failures += 0 if ok else 1
            del d_pid
            hh.peer_close()
            dist.barrier()
            hh.close()
            del d_mask
    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    h.peer_close()
    dist.barrier()  # nobody frees its exchange buffer while a peer still has it mapped
    h.close()
    dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)
    if rank == 0:
        print("MGPU PARITY OK world=%d" % world, flush=True)


if __name__ == "__main__":
    main()
