#!/bin/bash
# residency stamps on ONE GPU for several shapes
mkdir -p gpurun_out
for w in X_shard2_32768x16384_p8192 X_shard8_32768x4096_p2048 C5_32768x32768_p16384 C4_8192x8192_p4096; do
  DDC_DEBUG_TS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify --workload $w > gpurun_out/r2k_ts_1gpu_$w.json 2> gpurun_out/r2k_ts_1gpu_$w.log; echo "$w rc=$?"
  grep -a "ddc r0" gpurun_out/r2k_ts_1gpu_$w.log | tail -3 | cut -c1-500
done
