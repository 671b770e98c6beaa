#!/bin/bash
# 1 GPU: full suite on ABI v7 (in-kernel flags, emulated ranks)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run5_pytest.log
grep -v "Warning\|warn\|^  \|^$" gpurun_out/r2_run5_pytest.log | tail -15
