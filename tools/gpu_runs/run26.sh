#!/bin/bash
# 1 GPU, final build: full GPU test suite, smoke, default bench line, and the e2e loop at the 8-GPU per-rank size (host-side breakdown)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --pairs 4096 --steps 200 --no-parity --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_n1_4096_e2e.json 2> gpurun_out/r2_bench_n1_4096_e2e.err; echo "bench4096 rc=$?"
