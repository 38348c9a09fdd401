#!/bin/bash
# round 2, pass d (one B200): level-synchronous cut kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
timeout 300 python scripts/knob_sweep.py --workloads C2_528x522_p64,C3_4096x4096_p1024,C4_8192x8192_p4096,X_shard8_32768x4096_p2048,C5_32768x32768_p16384 \
  --sets "DDC_PDL=1" --ts > gpurun_out/r2d_sweep.jsonl 2> gpurun_out/r2d_sweep_ts.log; echo "sweep rc=$?"
NCU="ncu --set full --import-source on --clock-control none"
SW="python scripts/knob_sweep.py --sets DDC_PDL=0 --steps 1 --warmup 2"
timeout 300 $NCU -k regex:"k_xcuts|k_ycuts" -c 2 --launch-skip 4 -o gpurun_out/r2d_ncu_cuts_c5 $SW --workloads C5_32768x32768_p16384 > gpurun_out/r2d_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 python bench.py > gpurun_out/r2d_bench_c5_1gpu.json 2> gpurun_out/r2d_bench_c5_1gpu.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/r2d_sweep.jsonl
grep "ddc r0" gpurun_out/r2d_sweep_ts.log | awk 'NR%6==1 || NR%6==2' | cut -c1-420
tail -c 3000 gpurun_out/r2d_bench_c5_1gpu.json; tail -5 gpurun_out/r2d_bench_c5_1gpu.err
