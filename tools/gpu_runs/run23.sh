#!/bin/bash
# 1 GPU: configs[4] sweep with the round-2 build; ncu --set full of all five kernels of a step on the final build
set -x
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --out gpurun_out/r2_sweep_n1.json > gpurun_out/r2_sweep_n1.log 2>&1; echo "sweep rc=$?"; grep "^d=" gpurun_out/r2_sweep_n1.log | tail -42
ncu --set full --clock-control none --import-source on -k regex:"ntxent_tile|normalize_cast|dh_kernel|finalize_loss" -s 10 -c 5 -o gpurun_out/r2_step_full_final -f python tools/prof_step.py 32768 128 3 > gpurun_out/r2_run23_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_step_full_final.ncu-rep
