#!/bin/bash
mkdir -p gpurun_out
python scripts/knob_sweep.py --workloads C5_32768x32768_p16384,C2_528x522_p64,C3_4096x4096_p1024 --steps 30 --sets 'DDC_SIDE_PDL=0;DDC_SIDE_PDL=1;DDC_SIDE_PDL=0;DDC_SIDE_PDL=1' > gpurun_out/r2u_sweep.jsonl 2> gpurun_out/r2u_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2u_sweep.jsonl"):
    d = json.loads(l)
    print(d["workload"][:8], d["knobs"], d["ms_per_step"], d["best_ms"], d["same_result_as_first_set"])
PY
