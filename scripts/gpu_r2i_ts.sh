#!/bin/bash
# round 2: when do the blocks of each kernel of the chain get onto an SM, on N GPUs (DDC_DEBUG_TS residency stamps)
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29700
for e in 0 31; do
  port=$((port+1))
  DDC_EARLY=$e DDC_DEBUG_TS=1 timeout 600 $TR --master-port $port bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/r2i_ts_${N}gpu_early$e.json 2> gpurun_out/r2i_ts_${N}gpu_early$e.log; echo "ts rc=$?"
  grep -a "resident" gpurun_out/r2i_ts_${N}gpu_early$e.log | tail -4 | cut -c1-400
done
for e in 0 12; do
  DDC_EARLY=$e DDC_DEBUG_TS=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-verify --workload X_shard2_32768x16384_p8192 > gpurun_out/r2i_ts_1gpu_early$e.json 2> gpurun_out/r2i_ts_1gpu_early$e.log; echo "ts rc=$?"
  grep -a "ddc r0" gpurun_out/r2i_ts_1gpu_early$e.log | tail -3 | cut -c1-500
done
