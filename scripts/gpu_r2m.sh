#!/bin/bash
# round 2: per-block row flags, strip-K at N GPUs; optionally N independent single-GPU runs side by side
N=${1:-2}
T=${2:-r2m}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29800
timeout 900 $TR --master-port $port tests/mgpu_worker.py --big > gpurun_out/${T}_mgpu${N}_parity.log 2>&1; echo "parity rc=$?"; grep -E "MISMATCH|PARITY|Error|error" gpurun_out/${T}_mgpu${N}_parity.log | head -20
for knobs in ${KNOBS:-"DDC_ROW_FLAGS=1" "DDC_ROW_FLAGS=0" "DDC_STRIP_K=4" "DDC_STRIP_K=1" "DDC_EARLY=17" "DDC_ROW_FLAGS=1"}; do
  port=$((port+1))
  env $knobs timeout 600 $TR --master-port $port bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/${T}_bench_c5_${N}gpu_$knobs.json 2> gpurun_out/${T}_bench_c5_${N}gpu_$knobs.err; echo "$knobs rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_c5_${N}gpu_$knobs.json").read().strip().splitlines()[-1])
print("$knobs ms_per_step", d["ms_per_step"], "parity", d.get("parity",{}).get("passed"), d["roofline"]["stage_ms"])
PY
done
port=$((port+1))
DDC_DEBUG_TS=1 timeout 600 $TR --master-port $port bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ts_${N}gpu.json 2> gpurun_out/${T}_ts_${N}gpu.log; echo "ts rc=$?"
grep -a -o "ddc r0\] scan[^\[]*" gpurun_out/${T}_ts_${N}gpu.log | head -4 | cut -c1-420
if [ "$3" = "side" ]; then
  # N independent single-GPU decompositions of one shard's size, side by side: is a kernel slower because the box is busy?
  for g in $(seq 0 $((N-1))); do
    CUDA_VISIBLE_DEVICES=$g DDC_DEBUG_TS=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-verify --workload X_shard8_32768x4096_p2048 > gpurun_out/${T}_side_$g.json 2> gpurun_out/${T}_side_$g.log &
  done
  wait
  for g in 0 $((N-1)); do grep -a -o "ddc r0\] scan[^\[]*" gpurun_out/${T}_side_$g.log | head -4 | tail -2 | cut -c1-420; done
fi
