#!/bin/bash
mkdir -p gpurun_out
python scripts/knob_sweep.py --workloads X_shard8_32768x4096_p2048,C5_32768x32768_p16384 --steps 30 --sets 'DDC_X=0;DDC_X=1' > gpurun_out/r2last_sweep.jsonl 2>/dev/null; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2last_sweep.jsonl"):
    d = json.loads(l)
    print(d["workload"][:8], d["ms_per_step"], d["same_result_as_first_set"], d["stage_ms_profiled"])
PY
