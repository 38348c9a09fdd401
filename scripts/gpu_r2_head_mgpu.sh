#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tests/mgpu_worker.py --big > gpurun_out/r2yy_mgpu${N}_parity.log 2>&1; echo "parity rc=$?"; grep -E "MISMATCH|PARITY|Error|error" gpurun_out/r2yy_mgpu${N}_parity.log | head -5
timeout 300 $TR --master-port 29542 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/r2yy_bench_c5_${N}gpu.json 2> gpurun_out/r2yy_bench_c5_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2yy_bench_c5_${N}gpu.json").read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print("ms_per_step", round(d["ms_per_step"], 5), "parity", d["parity"]["passed"], "e2e", e.get("ms_per_step"), "same", e.get("same_result_as_c_abi"))
PY
