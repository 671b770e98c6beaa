#!/bin/bash
# 4 GPUs: final bench line
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29537"
timeout 200 $TR bench.py --gpus 4 --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/r2_bench_n4_final.err; echo "bench rc=$?"
