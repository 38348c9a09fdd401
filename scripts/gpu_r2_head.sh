#!/bin/bash
# HEAD sanity on one B200: GPU tests, smoke(), the bench lines
mkdir -p gpurun_out
T=${1:-r2zz}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in C5_32768x32768_p16384 C4_8192x8192_p4096 C3_4096x4096_p1024 C2_528x522_p64; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/${T}_bench_${w}_1gpu.json 2> gpurun_out/${T}_bench_${w}_1gpu.err; echo "bench $w rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench_${w}_1gpu.json").read().strip().splitlines()[-1])
print("$w ms_per_step", round(d["ms_per_step"],5), "parity", d["parity"]["passed"], "e2e", round(d["e2e"]["ms_per_step"],3), "roof", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"])
PY
done
