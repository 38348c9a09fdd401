#!/bin/bash
mkdir -p gpurun_out
python scripts/knob_sweep.py --workloads C2_528x522_p64,C3_4096x4096_p1024,C4_8192x8192_p4096,X_shard8_32768x4096_p2048,C5_32768x32768_p16384 --steps 30 --sets 'DDC_X=0;DDC_X=1;DDC_X=2' > gpurun_out/r2t_sweep.jsonl 2> gpurun_out/r2t_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2t_sweep.jsonl"):
    d = json.loads(l)
    print(d["workload"][:8], d["ms_per_step"], d["best_ms"], d["same_result_as_first_set"], d["stage_ms_profiled"])
PY
