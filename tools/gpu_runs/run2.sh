#!/bin/bash
set -x
mkdir -p gpurun_out
export MAAI_DEBUG_SEGV=1
timeout 300 python -X faulthandler -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "symmetric_forward" > gpurun_out/r2_run2_symfwd.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run2_symfwd.log
tail -40 gpurun_out/r2_run2_symfwd.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "symmetric_forward and 2-96" > gpurun_out/r2_run2_sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run2_sanitizer.log
grep -v "^$" gpurun_out/r2_run2_sanitizer.log | head -60
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multirank.py::test_emulated_ranks_symmetric_forward > gpurun_out/r2_run2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run2_pytest.log
tail -30 gpurun_out/r2_run2_pytest.log
V=multimodal-active-ai_b200/variants/pdl31.so
MAAI_PDL=1 MAAI_DEBUG_LIB=$V timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2_run2_pdl_pytest.log 2>&1; echo "pdl pytest rc=$?" >> gpurun_out/r2_run2_pdl_pytest.log
tail -15 gpurun_out/r2_run2_pdl_pytest.log
