#!/bin/bash
# round 2, pass b (one B200): GPU tests, strip-row variants, ncu source-level captures of the cut kernels, default bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
timeout 300 python scripts/knob_sweep.py --workloads C2_528x522_p64,C3_4096x4096_p1024,C4_8192x8192_p4096,X_shard8_32768x4096_p2048,C5_32768x32768_p16384 \
  --sets "DDC_PDL=1;DDC_STRIP_K=1;DDC_STRIP_K=2;DDC_STRIP_K=4;DDC_STRIP_K=8" --ts > gpurun_out/r2b_sweep.jsonl 2> gpurun_out/r2b_sweep_ts.log; echo "sweep rc=$?"
NCU="ncu --set full --import-source on --clock-control none"
SW="python scripts/knob_sweep.py --sets DDC_PDL=0 --steps 1 --warmup 2"
timeout 300 $NCU -k regex:k_xcuts -c 1 --launch-skip 2 -o gpurun_out/r2b_ncu_xcuts_c5 $SW --workloads C5_32768x32768_p16384 > gpurun_out/r2b_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 $NCU -k regex:k_ycuts -c 1 --launch-skip 2 -o gpurun_out/r2b_ncu_ycuts_c5 $SW --workloads C5_32768x32768_p16384 > gpurun_out/r2b_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 300 $NCU -k regex:"k_strip_rows|k_scan_mask|k_label" -c 3 --launch-skip 6 -o gpurun_out/r2b_ncu_stream_shard8 $SW --workloads X_shard8_32768x4096_p2048 > gpurun_out/r2b_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 300 $NCU -k regex:"k_xcuts|k_ycuts" -c 2 --launch-skip 4 -o gpurun_out/r2b_ncu_cuts_c2 $SW --workloads C2_528x522_p64 > gpurun_out/r2b_ncu4.log 2>&1; echo "ncu4 rc=$?"
timeout 900 python bench.py > gpurun_out/r2b_bench_c5_1gpu.json 2> gpurun_out/r2b_bench_c5_1gpu.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/r2b_sweep.jsonl
tail -c 2500 gpurun_out/r2b_bench_c5_1gpu.json; tail -5 gpurun_out/r2b_bench_c5_1gpu.err
ls -la gpurun_out/*.ncu-rep
