#!/bin/bash
# 8 GPUs: final bench line (timed run through the C++ binding, parity gate on)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err; echo "bench rc=$?"
