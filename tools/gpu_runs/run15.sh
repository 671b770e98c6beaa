#!/bin/bash
# 8 GPUs: parity worker incl. the direct symmetric forward, bench default vs direct vs staged
set -x
mkdir -p gpurun_out
export MAAI_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n8_flags.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n8_flags.log
grep "DIST_\|rc=\|modes\|convergence\|chained\|raised" gpurun_out/r2_dist_n8_flags.log; grep -i "timeout" gpurun_out/r2_dist_n8_flags.log | head -3
B="bench.py --gpus 8 --steps 20 --warmup 5 --require-peer --no-secondary"
for cfg in "default:" "direct:MAAI_FWD_SYM_MULTI=direct" "staged:MAAI_FWD_SYM_MULTI=staged"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 $TR $B > gpurun_out/r2_bench_n8_$name.json 2> gpurun_out/r2_bench_n8_$name.err
  echo "== $name rc=$?"; grep "^\[rank [07]\]" gpurun_out/r2_bench_n8_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_n8_$name.json')); c=d['config']
    print('   ms', round(d['ms_per_step'],4), 'median', round(c['ms_median'],4), 'frac', round(c['step_frac_bf16_peak'],3), 'sustained', round(c['run_lengths']['sustained']['ms_per_step'],4), 'host', round(c['host_issue_ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['schedule'], 'copies', round(d['e2e']['copies_only_ms_per_step'],4), 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'], d['parity']['loss_rel'], c.get('sym_forward_mode'), 'launches', d['gpu_launches'], d['clocks'])
except Exception as e: print('   failed', e)
PY
done
