#!/bin/bash
mkdir -p gpurun_out
python scripts/knob_sweep.py --workloads C2_528x522_p64,C3_4096x4096_p1024 --steps 40 --sets 'DDC_GATE=1;DDC_GATE=0;DDC_GATE=1 DDC_LABEL_RPC=8;DDC_GATE=0 DDC_LABEL_RPC=8;DDC_GATE=1 DDC_LABEL_RPC=16;DDC_GATE=0 DDC_LABEL_RPC=16;DDC_GATE=1;DDC_GATE=0;DDC_GATE=1 DDC_LABEL_RPC=8;DDC_GATE=0 DDC_LABEL_RPC=8' > gpurun_out/r2w_sweep.jsonl 2> gpurun_out/r2w_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2w_sweep.jsonl"):
    d = json.loads(l)
    print(d["workload"][:3], d["knobs"], d["ms_per_step"], d["best_ms"], d["same_result_as_first_set"])
PY
