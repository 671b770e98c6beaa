#!/bin/bash
# GPU call 1 of round 2: parity suite on the new ABI-v6 path, PDL variant check, small-batch timing
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run1_pytest.log
tail -5 gpurun_out/r2_run1_pytest.log
V=multimodal-active-ai_b200/variants/pdl31.so
MAAI_PDL=1 MAAI_DEBUG_LIB=$V timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_run1_pdl_pytest.log 2>&1; echo "pdl pytest rc=$?" >> gpurun_out/r2_run1_pdl_pytest.log
tail -5 gpurun_out/r2_run1_pdl_pytest.log
{
python tools/small_batch_time.py 256 4096
MAAI_PDL=1 MAAI_DEBUG_LIB=$V python tools/small_batch_time.py 256 4096
python tools/quick_time.py 32768 128 30
MAAI_PDL=1 MAAI_DEBUG_LIB=$V python tools/quick_time.py 32768 128 30
python tools/quick_time.py 4096 128 100
MAAI_PDL=1 MAAI_DEBUG_LIB=$V python tools/quick_time.py 4096 128 100
} > gpurun_out/r2_run1_timing.log 2>&1
cat gpurun_out/r2_run1_timing.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_run1_bench.json
tail -5 gpurun_out/r2_run1_bench.err
