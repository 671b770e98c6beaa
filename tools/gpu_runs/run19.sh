#!/bin/bash
# 2 GPUs: non-current-device test, config-4 convergence (1 GPU and 2 GPUs), chained-view K1 span
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "non_current or two_gpu" > gpurun_out/r2_run19_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_run19_pytest.log
timeout 900 python tools/config4_convergence.py --steps 100 --batch-per-gpu 128 --out gpurun_out/r2_config4_convergence_n1.json > gpurun_out/r2_run19_conv1.log 2>&1; echo "conv1 rc=$?"; tail -13 gpurun_out/r2_run19_conv1.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR tools/config4_convergence.py --steps 60 --batch-per-gpu 128 --out gpurun_out/r2_config4_convergence_n2.json > gpurun_out/r2_run19_conv2.log 2>&1; echo "conv2 rc=$?"; grep "^step\|max_rel" gpurun_out/r2_run19_conv2.log | tail -8 | cut -c1-300
timeout 600 $TR tools/chain_time.py 32768 > gpurun_out/r2_run19_chain_n2.log 2>&1; grep "^W=" gpurun_out/r2_run19_chain_n2.log
