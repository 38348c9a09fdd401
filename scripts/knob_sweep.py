#!/usr/bin/env python
"""One GPU, one process: time ddc_partition (device-resident mask, boxes + pid + neighbour tables) for several
workloads under several settings of the library's environment knobs (read by ddc_create), so that one gpurun
call answers "what did this change buy".  Prints one JSON line per (workload, knobs).

    python scripts/knob_sweep.py [--workloads A,B] [--sets 'K1=V1 K2=V2;K3=V3'] [--steps 20] [--ts]

Timing as in bench.py: CUDA events on the launching stream, warm-up first, L2 flushed between iterations when the
mask and the pid map of a rank could sit in it.  --ts: one more step per setting with DDC_DEBUG_TS=1 (the stamps
go to stderr).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import L2_BYTES, WORKLOADS  # noqa: E402
from domain_decomp_b200 import capi, verify  # noqa: E402

DEFAULT_SETS = ";".join([
    "DDC_PDL=0 DDC_WARM=0 DDC_FUSE_FIN=0",
    "DDC_PDL=1 DDC_WARM=0 DDC_FUSE_FIN=0",
    "DDC_PDL=1 DDC_WARM=1 DDC_FUSE_FIN=0",
    "DDC_PDL=1 DDC_WARM=0 DDC_FUSE_FIN=1",
    "DDC_PDL=1 DDC_WARM=1 DDC_FUSE_FIN=1",
])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="C2_528x522_p64,C3_4096x4096_p1024,C4_8192x8192_p4096,X_shard8_32768x4096_p2048")
    ap.add_argument("--sets", default=DEFAULT_SETS)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--ts", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flags = capi.WANT_PID | capi.WANT_NEIGHBOURS
    for wl in args.workloads.split(","):
        nx, ny, P, land, seed, px, py = WORKLOADS[wl]
        d_mask = torch.empty((ny, nx), dtype=torch.int32, device=dev)
        need_flush = ny * nx * 8 < 2 * L2_BYTES
        ref_digest = None
        for kset in args.sets.split(";"):
            env = dict(kv.split("=") for kv in kset.split())
            for k in [k for k in os.environ if k.startswith("DDC_")]:
                del os.environ[k]
            os.environ.update(env)
            h = capi.Handle(0)
            h.set_stream(stream.cuda_stream)
            h.generate_mask_device(d_mask.data_ptr(), nx, ny, seed, land)
            h.set_mask_device(d_mask.data_ptr(), nx, ny)
            for _ in range(args.warmup):
                h.partition(P, px, py, flags)
            torch.cuda.synchronize()
            times = []
            if need_flush:
                for _ in range(args.steps):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    h.partition(P, px, py, flags | capi.ASYNC)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1))
                ms = sum(times) / len(times)
                best = min(times)
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    h.partition(P, px, py, flags | capi.ASYNC)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = best = e0.elapsed_time(e1) / args.steps
            h.partition(P, px, py, flags | capi.PROFILE)
            st = h.stats()
            digest = verify.handle_digest(h)
            ref_digest = ref_digest or digest
            print(json.dumps({"workload": wl, "knobs": env, "ms_per_step": round(ms, 5), "best_ms": round(best, 5),
                              "launches": st["gpu_launches"], "same_result_as_first_set": digest == ref_digest,
                              "stage_ms_profiled": {k: round(v, 4) for k, v in st["stage_ms"].items()}}), flush=True)
            h.close()
            if args.ts:
                os.environ["DDC_DEBUG_TS"] = "1"
                h = capi.Handle(0)
                h.set_stream(stream.cuda_stream)
                h.set_mask_device(d_mask.data_ptr(), nx, ny)
                sys.stderr.write("== %s %s\n" % (wl, kset))
                for _ in range(3):
                    flush.fill_(1)
                    h.partition(P, px, py, flags)
                h.close()
        del d_mask


if __name__ == "__main__":
    main()
