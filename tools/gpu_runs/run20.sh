#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/config4_convergence.py --steps 100 --batch-per-gpu 128 --out gpurun_out/r2_config4_convergence_n1.json > gpurun_out/r2_run20_conv1.log 2>&1; echo "conv1 rc=$?"; tail -13 gpurun_out/r2_run20_conv1.log | cut -c1-600
