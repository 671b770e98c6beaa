#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n2_flags.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n2_flags.log
grep "DIST_\|rc=\|convergence" gpurun_out/r2_dist_n2_flags.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --require-peer --no-secondary > gpurun_out/r2_run13_n2.json 2> gpurun_out/r2_run13_n2.err
echo "== n2 rc=$?"; grep "^\[rank" gpurun_out/r2_run13_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_run13_n2.json')); c=d['config']; print('   ms', round(d['ms_per_step'],4), 'median', round(c['ms_median'],4), 'clocks', d['clocks'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['schedule'], d['e2e'].get('copies_only_ms_per_step'), d['e2e']['serial_ms_per_step'])"
timeout 900 python tools/simclr_step.py --steps 10 --out gpurun_out/r2_simclr_step_n1.json > gpurun_out/r2_simclr_step_n1.log 2>&1; echo "simclr rc=$?"; tail -12 gpurun_out/r2_simclr_step_n1.log
