#!/bin/bash
mkdir -p gpurun_out
./scripts/median_probe > gpurun_out/r2f_median_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2f_median_probe.txt
