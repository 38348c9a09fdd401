#!/bin/bash
# round 2, definitive multi-GPU pass with the committed defaults: parity at BASELINE sizes, the bench line (parity block,
# plugin e2e), stamps, the NCCL fallback and C4 beside it, one variant of the second stream's priority
N=${1:-8}
T=${2:-r2z}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tests/mgpu_worker.py --big > gpurun_out/${T}_mgpu${N}_parity.log 2>&1; echo "parity rc=$?"; grep -E "MISMATCH|PARITY|Error|error" gpurun_out/${T}_mgpu${N}_parity.log | head -20
timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${T}_bench_c5_${N}gpu.json 2> gpurun_out/${T}_bench_c5_${N}gpu.err; echo "bench rc=$?"
DDC_SIDE_PRIO=0 timeout 300 $TR --master-port 29546 bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/${T}_bench_c5_${N}gpu_sideprio0.json 2>/dev/null; echo "side prio 0 rc=$?"
timeout 300 $TR --master-port 29547 bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/${T}_bench_c5_${N}gpu_again.json 2>/dev/null; echo "again rc=$?"
DDC_DEBUG_TS=1 timeout 300 $TR --master-port 29543 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ts_${N}gpu.json 2> gpurun_out/${T}_ts_${N}gpu.log; echo "ts rc=$?"
timeout 300 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu --exchange nccl > gpurun_out/${T}_bench_c5_${N}gpu_nccl.json 2> gpurun_out/${T}_bench_c5_${N}gpu_nccl.err; echo "nccl bench rc=$?"
timeout 300 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu --workload C4_8192x8192_p4096 > gpurun_out/${T}_bench_c4_${N}gpu.json 2> gpurun_out/${T}_bench_c4_${N}gpu.err; echo "c4 bench rc=$?"
python - <<PY
import json
for f in ("bench_c5_${N}gpu", "bench_c5_${N}gpu_sideprio0", "bench_c5_${N}gpu_again", "bench_c5_${N}gpu_nccl", "bench_c4_${N}gpu"):
    try:
        d = json.loads(open("gpurun_out/${T}_%s.json" % f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f, "ms_per_step", round(d["ms_per_step"], 5), "parity", (d.get("parity") or {}).get("passed"), "frac", d["roofline"]["pipeline"]["frac_of_aggregate_peak"],
              "e2e ms", e.get("ms_per_step"), d["roofline"]["stage_ms"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
grep -a -o "ddc r0\] scan[^\[]*" gpurun_out/${T}_ts_${N}gpu.log | head -4 | cut -c1-420
