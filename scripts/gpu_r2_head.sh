#!/bin/bash
# HEAD sanity on one B200: GPU tests, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2zz_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2zz_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2zz_bench_c5_1gpu.json 2> gpurun_out/r2zz_bench_c5_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2zz_bench_c5_1gpu.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "parity", d["parity"]["passed"], "e2e", d["e2e"]["ms_per_step"], "roof", d["roofline"]["frac"], "launches", d["gpu_launches"])
PY
