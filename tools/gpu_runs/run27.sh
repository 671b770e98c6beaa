#!/bin/bash
# 1 GPU: span brackets inside the C++ binding (bench main run goes through the binding now): parity file, bench lines
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_20_extspans.json 2> gpurun_out/r2_bench_n1_20_extspans.err; echo "bench rc=$?"
timeout 200 python bench.py --pairs 4096 --steps 200 --no-parity --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_n1_4096_extspans.json 2> gpurun_out/r2_bench_n1_4096_extspans.err; echo "bench4096 rc=$?"
python tools/quick_time.py 32768 128 20
