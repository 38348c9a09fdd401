#!/bin/bash
# what is different about this box? (kernel-to-kernel start latency, environment, driver state)
mkdir -p gpurun_out
T=${1:-1gpu}
{
echo "== env"; env | grep -E 'CUDA|NCCL|NVIDIA|OMP|TORCH' | sort
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name,persistence_mode,compute_mode,mig.mode.current,clocks.sm,clocks.max.sm,power.limit,ecc.mode.current --format=csv
nvidia-smi -q -i 0 | grep -iE 'Addressing|Virtualization|Fabric|State|Status|MIG|Persistence|Confidential|GSP|Operation' | head -30
nvidia-smi topo -m 2>/dev/null | head -14
echo "== cpu"; nproc; grep -m1 'model name' /proc/cpuinfo; cat /proc/loadavg
echo "== pdl probe"; ./scripts/pdl_probe
} > gpurun_out/r2j_box_$T.txt 2>&1
tail -25 gpurun_out/r2j_box_$T.txt
