#!/bin/bash
mkdir -p gpurun_out
export DDC_GATE=0
for w in C2_528x522_p64 C4_8192x8192_p4096; do
  timeout 300 ncu --set full --import-source on --warp-sampling-interval 0 --clock-control none -k regex:"k_ycuts|k_xcuts" --launch-skip 12 -c 2 \
    -o gpurun_out/r2x_ncu_cuts_$w python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify --workload $w > gpurun_out/r2x_ncu_$w.log 2>&1; echo "ncu $w rc=$?"
done
ls -la gpurun_out/r2x_*
