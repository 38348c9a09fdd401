#!/bin/bash
mkdir -p gpurun_out
python scripts/knob_sweep.py --workloads C2_528x522_p64,C3_4096x4096_p1024,C4_8192x8192_p4096 --steps 40 --sets 'DDC_X=0;DDC_STRIP_K=1;DDC_STRIP_K=2;DDC_STRIP_K=4;DDC_STRIP_K=8;DDC_SCAN_TAIL=0;DDC_SCAN_TAIL=50;DDC_EARLY=25;DDC_EARLY=29;DDC_X=1' > gpurun_out/r2w2_sweep.jsonl 2> gpurun_out/r2w2_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2w2_sweep.jsonl"):
    d = json.loads(l)
    print(d["workload"][:3], d["knobs"], d["ms_per_step"], d["best_ms"], d["same_result_as_first_set"], "rows", d["stage_ms_profiled"]["strip_rows"], "scan", d["stage_ms_profiled"]["mask_scan"])
PY
