#!/bin/bash
# round 2, pass a (one B200): GPU tests, knob sweep with time stamps, shard-size tuning, the default bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest.log
timeout 400 python scripts/knob_sweep.py --ts > gpurun_out/r2a_sweep.jsonl 2> gpurun_out/r2a_sweep_ts.log; echo "sweep rc=$?"
timeout 300 python scripts/knob_sweep.py --workloads X_shard8_32768x4096_p2048 \
  --sets "DDC_SCAN_RPC=32;DDC_SCAN_RPC=64;DDC_SCAN_RPC=128;DDC_SCAN_RPC=256;DDC_LABEL_RPC=16;DDC_LABEL_RPC=64;DDC_LABEL_RPC=128" \
  > gpurun_out/r2a_sweep_rpc.jsonl 2> gpurun_out/r2a_sweep_rpc.err; echo "rpc sweep rc=$?"
timeout 400 python scripts/knob_sweep.py --workloads C5_32768x32768_p16384 --sets "DDC_PDL=0 DDC_WARM=0 DDC_FUSE_FIN=0;DDC_PDL=1 DDC_WARM=1 DDC_FUSE_FIN=1" --ts \
  > gpurun_out/r2a_sweep_c5.jsonl 2> gpurun_out/r2a_sweep_c5_ts.log; echo "c5 sweep rc=$?"
timeout 600 python bench.py > gpurun_out/r2a_bench_c5_1gpu.json 2> gpurun_out/r2a_bench_c5_1gpu.err; echo "bench rc=$?"
cat gpurun_out/r2a_sweep.jsonl gpurun_out/r2a_sweep_rpc.jsonl gpurun_out/r2a_sweep_c5.jsonl | cut -c1-400
tail -c 1500 gpurun_out/r2a_bench_c5_1gpu.json
