#!/bin/bash
# 2 GPUs: multi-rank step through the C++ binding (ntxent_loss_peer): parity worker, then bench with the binding on / off
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tests/_dist_gpu_worker.py gpurun_out/r2_dist_n2_ext.json > gpurun_out/r2_dist_n2_ext.log 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/r2_dist_n2_ext.log
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n2_ext.json 2> gpurun_out/r2_bench_n2_ext.err; echo "bench ext rc=$?"; cat gpurun_out/r2_bench_n2_ext.json
MAAI_FAST_EXT=0 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --require-peer --no-parity > gpurun_out/r2_bench_n2_noext.json 2> gpurun_out/r2_bench_n2_noext.err; echo "bench noext rc=$?"; cat gpurun_out/r2_bench_n2_noext.json
