#!/bin/bash
# round 2, final pass on one B200: GPU tests, the bench lines of every BASELINE workload, the reference arm on the whole
# C5 workload, ncu launch list + full captures of the step's kernels, the CLI at C4
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r2z}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_c5_1gpu.json 2> gpurun_out/${T}_bench_c5_1gpu.err; echo "bench c5 rc=$?"
for W in C4_8192x8192_p4096 C3_4096x4096_p1024 C2_528x522_p64; do
  timeout 600 python bench.py --workload $W > gpurun_out/${T}_bench_${W}_1gpu.json 2> gpurun_out/${T}_bench_${W}_1gpu.err; echo "bench $W rc=$?"
done
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference_c5.json 2> gpurun_out/${T}_bench_reference_c5.err; echo "reference rc=$?"
# ncu: kernels are serialised under the profiler, so the device-side gate of the second stream is switched off
export DDC_GATE=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c5_1gpu.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_scan_mask|k_xcuts|k_strip_rows|k_ycuts|k_label" --launch-skip 15 -c 5 \
  -o gpurun_out/${T}_ncu_full_c5_1gpu python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
unset DDC_GATE
# the CLI at C4: a classic netCDF grid written with scipy, decomp --parts 4096 --stats
python - <<'PY' > gpurun_out/${T}_cli_c4.log 2>&1
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from scipy.io import netcdf_file
from oracle import oracle as orc
nx = ny = 8192
m = orc.generate_mask(nx, ny, 1, 0.60)
os.makedirs("/tmp/cli", exist_ok=True)
f = netcdf_file("/tmp/cli/c4.nc", "w", version=2)
f.createDimension("x", nx); f.createDimension("y", ny)
f.createVariable("mask", "i4", ("y", "x"))[:] = m
f.close()
for extra in ([], ["--gpus", "1"]):
    t = time.time()
    out = subprocess.run([os.path.join(os.getcwd(), "domain_decomp_b200", "decomp"), "-g", "c4.nc", "--parts", "4096", "--stats"] + extra,
                         cwd="/tmp/cli", capture_output=True, text=True)
    print("rc", out.returncode, "wall %.3f s" % (time.time() - t)); print(out.stdout); print(out.stderr[-500:])
PY
echo "cli rc=$?"; head -12 gpurun_out/${T}_cli_c4.log
for f in gpurun_out/${T}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f'))
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, 'e2e', (d.get('e2e') or {}).get('ms_per_step'), (d.get('e2e') or {}).get('path','')[:40], 'parity', (d.get('parity') or {}).get('passed'), 'roof', (d.get('roofline') or {}).get('frac'), 'pipe', ((d.get('roofline') or {}).get('pipeline') or {}).get('frac_of_aggregate_peak'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('ERR', e)
"; done
ls -la gpurun_out/${T}_*
