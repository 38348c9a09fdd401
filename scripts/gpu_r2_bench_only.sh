#!/bin/bash
# one bench line on N GPUs with the committed defaults (parity block, plugin e2e)
N=${1:-8}
T=${2:-r2z}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${T}_bench_c5_${N}gpu.json 2> gpurun_out/${T}_bench_c5_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench_c5_${N}gpu.json").read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print("ms_per_step", round(d["ms_per_step"], 5), "parity", (d.get("parity") or {}).get("passed"), "frac", d["roofline"]["pipeline"]["frac_of_aggregate_peak"], "e2e ms", e.get("ms_per_step"), "clocks", d.get("clocks"))
PY
