#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
for j in 1 0 1; do
  for w in C5_32768x32768_p16384 C4_8192x8192_p4096 C3_4096x4096_p1024 C2_528x522_p64 X_shard8_32768x4096_p2048; do
    DDC_DEV_JOIN=$j timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $w > gpurun_out/r2q_bench_${w}_join$j.json 2>/dev/null
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_bench_${w}_join$j.json").read().strip().splitlines()[-1])
print("$w dev_join=$j ms_per_step", round(d["ms_per_step"],5), "parity", d.get("parity",{}).get("passed"))
PY
  done
done
