#!/bin/bash
# 2 GPUs: real-rank parity (flags and barrier modes), workspace reuse, chained views; bench N=2
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n2_flags.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n2_flags.log
grep -v "Warning\|warn" gpurun_out/r2_dist_n2_flags.log | tail -25
MAAI_PEER_FLAGS=0 timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n2_barriers.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n2_barriers.log
grep -v "Warning\|warn" gpurun_out/r2_dist_n2_barriers.log | tail -8
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n2_flags.json 2> gpurun_out/r2_bench_n2_flags.err; echo "bench rc=$?"
MAAI_PEER_FLAGS=0 timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n2_barriers.json 2> gpurun_out/r2_bench_n2_barriers.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ('flags','barriers'):
    try:
        d=json.load(open(f'gpurun_out/r2_bench_n2_{n}.json'))
        print(n, d['ms_per_step'], d['config']['gather_mode'], d['config']['sym_forward'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['schedule'], 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'], d['parity']['loss_rel'])
        print('   spans', {k:round(v,4) for k,v in d['config']['span_ms_mean'].items() if v})
    except Exception as e: print(n, 'failed', e)
PY
tail -5 gpurun_out/r2_bench_n2_flags.err
