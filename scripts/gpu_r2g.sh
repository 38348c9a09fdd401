#!/bin/bash
# round 2, pass g (one B200): gate + programmatic label launch, median probe
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
./scripts/median_probe > gpurun_out/r2g_median_probe.txt 2>&1; echo "probe rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.log
timeout 300 python scripts/knob_sweep.py --workloads C2_528x522_p64,C3_4096x4096_p1024,C4_8192x8192_p4096,X_shard8_32768x4096_p2048,C5_32768x32768_p16384 \
  --sets "DDC_GATE=0;DDC_GATE=1" --ts > gpurun_out/r2g_sweep.jsonl 2> gpurun_out/r2g_sweep_ts.log; echo "sweep rc=$?"
timeout 900 python bench.py > gpurun_out/r2g_bench_c5_1gpu.json 2> gpurun_out/r2g_bench_c5_1gpu.err; echo "bench rc=$?"
cat gpurun_out/r2g_median_probe.txt | cut -c1-300
cut -c1-200 gpurun_out/r2g_sweep.jsonl
grep "ddc r0" gpurun_out/r2g_sweep_ts.log | awk 'NR%6==1' | cut -c1-420
tail -c 3500 gpurun_out/r2g_bench_c5_1gpu.json; tail -5 gpurun_out/r2g_bench_c5_1gpu.err
