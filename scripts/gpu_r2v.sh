#!/bin/bash
# same box, two builds of the library: did the step get slower between two commits?
mkdir -p gpurun_out
for v in old new old new; do
  cp gpurun_ab/libddc_cuda_$v.so domain_decomp_b200/libddc_cuda.so
  python scripts/knob_sweep.py --workloads C5_32768x32768_p16384 --steps 30 --sets 'DDC_X=0' 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$v', d['workload'][:8], d['ms_per_step'], d['stage_ms_profiled'])
"
done
cp gpurun_ab/libddc_cuda_new.so domain_decomp_b200/libddc_cuda.so
