#!/bin/bash
mkdir -p gpurun_out
python scripts/knob_sweep.py --workloads C2_528x522_p64,C3_4096x4096_p1024,C4_8192x8192_p4096 --steps 30 --sets 'DDC_SCAN_RPC=0;DDC_SCAN_RPC=8;DDC_SCAN_RPC=16;DDC_SCAN_RPC=32;DDC_SCAN_RPC=64;DDC_SCAN_RPC=128;DDC_LABEL_RPC=8;DDC_LABEL_RPC=16;DDC_LABEL_RPC=64;DDC_SCAN_RPC=32 DDC_LABEL_RPC=16;DDC_SCAN_RPC=0' > gpurun_out/r2r_sweep.jsonl 2> gpurun_out/r2r_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2r_sweep.jsonl"):
    d = json.loads(l)
    print(d["workload"][:3], d["knobs"], d["ms_per_step"], d["best_ms"], d["same_result_as_first_set"], "scan", d["stage_ms_profiled"]["mask_scan"], "label", d["stage_ms_profiled"]["label"])
PY
