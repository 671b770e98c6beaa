#!/bin/bash
# 2 GPUs, final build: parity worker (every dataflow, buffer-set reuse, convergence, chained views) + bench line
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536"
timeout 600 $TR tests/_dist_gpu_worker.py gpurun_out/r2_dist_n2_final.json > gpurun_out/r2_dist_n2_final.log 2>&1; echo "dist rc=$?"; tail -4 gpurun_out/r2_dist_n2_final.log | cut -c1-250
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; echo "bench rc=$?"
