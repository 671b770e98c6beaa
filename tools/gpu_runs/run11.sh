#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run11_pytest.log
grep -v "Warning\|warn\|^  \|^$" gpurun_out/r2_run11_pytest.log | tail -5
{
python tools/ab_variants.py run 32768 128 30
MAAI_FWD_SYM=0 python tools/ab_variants.py run 32768 128 30
python tools/ab_variants.py run 32768 64 30
} > gpurun_out/r2_run11_ab.log 2>&1
grep "B=" gpurun_out/r2_run11_ab.log
