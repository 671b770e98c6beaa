#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --require-peer --no-secondary > gpurun_out/r2_run16_n2.json 2> gpurun_out/r2_run16_n2.err
echo "== n2 rc=$?"; grep "^\[rank" gpurun_out/r2_run16_n2.err; tail -3 gpurun_out/r2_run16_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run16_n2.json')); c=d['config']; e=d['e2e']
print('   ms', round(d['ms_per_step'],4), 'median', round(c['ms_median'],4), 'max', round(c['ms_max'],4), 'host', round(c['host_issue_ms_per_step'],4))
print('   e2e', round(e['ms_per_step'],4), e['schedule'], 'pipelined', round(e['pipelined_ms_per_step'],4), 'serial', round(e['serial_ms_per_step'],4), 'copies', round(e['copies_only_ms_per_step'],4), 'compute span', e['pipelined_compute_span_ms'], 'host', e['pipelined_host_issue_ms_per_step'])
PY
