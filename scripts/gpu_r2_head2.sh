#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2yy}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu.log
for w in C2_528x522_p64 C3_4096x4096_p1024 C4_8192x8192_p4096; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/${T}_bench_${w}_1gpu.json 2> gpurun_out/${T}_bench_${w}_1gpu.err; echo "bench $w rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench_${w}_1gpu.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("$w ms_per_step", round(d["ms_per_step"],5), "parity", d["parity"]["passed"], "plugin e2e", round(e["ms_per_step"],3), "c_abi e2e", round((e.get("c_abi") or {}).get("ms_per_step", 0),3), "same", e.get("same_result_as_c_abi"))
PY
done
