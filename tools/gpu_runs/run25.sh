#!/bin/bash
# 8 GPUs: bench with the C++ multi-rank binding (parity gate on), then the same with the binding off (e2e / host issue A/B)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n8_ext.json 2> gpurun_out/r2_bench_n8_ext.err; echo "bench ext rc=$?"
MAAI_FAST_EXT=0 timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --require-peer --no-parity > gpurun_out/r2_bench_n8_noext.json 2> gpurun_out/r2_bench_n8_noext.err; echo "bench noext rc=$?"
