#!/bin/bash
mkdir -p gpurun_out
for w in C4_8192x8192_p4096 C3_4096x4096_p1024 C2_528x522_p64; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/r2s_bench_${w}_1gpu.json 2> gpurun_out/r2s_bench_${w}_1gpu.err; echo "bench $w rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2s_bench_${w}_1gpu.json").read().strip().splitlines()[-1])
print("$w ms_per_step", round(d["ms_per_step"],5), "parity", d["parity"]["passed"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"])
PY
done
timeout 600 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -2
