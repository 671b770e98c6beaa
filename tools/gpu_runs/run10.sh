#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_run10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run10_pytest.log
grep -v "Warning\|warn\|^  \|^$" gpurun_out/r2_run10_pytest.log | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n2_flags.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n2_flags.log
grep "DIST_\|rc=" gpurun_out/r2_dist_n2_flags.log
B="bench.py --gpus 2 --steps 40 --warmup 5 --require-peer --no-secondary"
env timeout 600 $TR $B > gpurun_out/r2_run10_flags.json 2> gpurun_out/r2_run10_flags.err
echo "== flags rc=$?"; grep "^\[rank" gpurun_out/r2_run10_flags.err
python -c "
import json; d=json.load(open('gpurun_out/r2_run10_flags.json')); c=d['config']; print('   ms', round(d['ms_per_step'],4), 'median', round(c['ms_median'],4), 'clocks', d['clocks'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('copies_only_ms_per_step'))"
