#!/bin/bash
# round 2: which kernel boundaries of the chain are worth replacing by polled words (DDC_EARLY bit mask), on N GPUs
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
DDC_EARLY=31 timeout 900 $TR --master-port $port tests/mgpu_worker.py --big > gpurun_out/r2h_mgpu${N}_parity_early31.log 2>&1; echo "parity rc=$?"; grep -E "MISMATCH|PARITY|Error|error" gpurun_out/r2h_mgpu${N}_parity_early31.log | head -20
for e in 0 1 3 7 15 31 0 31; do
  port=$((port+1))
  DDC_EARLY=$e timeout 600 $TR --master-port $port bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2h_bench_c5_${N}gpu_early$e.json 2> gpurun_out/r2h_bench_c5_${N}gpu_early$e.err; echo "early=$e rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_bench_c5_${N}gpu_early$e.json").read().strip().splitlines()[-1])
print("early=$e ms_per_step", d["ms_per_step"], "parity", d.get("parity",{}).get("passed"))
PY
done
for e in 0 31; do
  port=$((port+1))
  DDC_EARLY=$e DDC_DEBUG_TS=1 timeout 600 $TR --master-port $port bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/r2h_ts_${N}gpu_early$e.json 2> gpurun_out/r2h_ts_${N}gpu_early$e.log; echo "ts rc=$?"
  grep "ddc r0\]" gpurun_out/r2h_ts_${N}gpu_early$e.log | grep scan | tail -2 | cut -c1-420
done
# one GPU: words 4 (label), 8 (rows) only
for e in 0 12; do
  for w in C5_32768x32768_p16384 C4_8192x8192_p4096 C3_4096x4096_p1024 C2_528x522_p64; do
    DDC_EARLY=$e timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $w > gpurun_out/r2h_bench_${w}_1gpu_early$e.json 2>/dev/null
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_bench_${w}_1gpu_early$e.json").read().strip().splitlines()[-1])
print("1 GPU $w early=$e ms_per_step", d["ms_per_step"], "parity", d.get("parity",{}).get("passed"))
PY
  done
done
