#!/bin/bash
# round 2, multi-GPU pass: parity of the row-sharded path (both exchange modes, BASELINE sizes through the golden
# digests), the bench line with its parity block and the plugin e2e (--gpus N threads), time stamps of every rank
N=${1:-2}
T=${2:-r2e}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 tests/mgpu_worker.py --big > gpurun_out/${T}_mgpu${N}_parity.log 2>&1; echo "parity rc=$?"; grep -E "MISMATCH|PARITY|Error|error" gpurun_out/${T}_mgpu${N}_parity.log | head -20
timeout 900 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_c5_${N}gpu.json 2> gpurun_out/${T}_bench_c5_${N}gpu.err; echo "bench rc=$?"
DDC_DEBUG_TS=1 timeout 600 $TR --master-port 29543 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ts_${N}gpu.json 2> gpurun_out/${T}_ts_${N}gpu.log; echo "ts rc=$?"
timeout 600 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu --exchange nccl > gpurun_out/${T}_bench_c5_${N}gpu_nccl.json 2> gpurun_out/${T}_bench_c5_${N}gpu_nccl.err; echo "nccl bench rc=$?"
timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu --workload C4_8192x8192_p4096 > gpurun_out/${T}_bench_c4_${N}gpu.json 2> gpurun_out/${T}_bench_c4_${N}gpu.err; echo "c4 bench rc=$?"
tail -c 4000 gpurun_out/${T}_bench_c5_${N}gpu.json; tail -3 gpurun_out/${T}_bench_c5_${N}gpu.err
grep -a -o "ddc r0\] scan[^\[]*" gpurun_out/${T}_ts_${N}gpu.log | head -4 | cut -c1-420
cut -c1-300 gpurun_out/${T}_bench_c5_${N}gpu_nccl.json gpurun_out/${T}_bench_c4_${N}gpu.json
