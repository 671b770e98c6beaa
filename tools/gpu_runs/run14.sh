#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tests/_emulated_direct_worker.py > gpurun_out/r2_run14_emu.log 2>&1; echo "emu rc=$?"; tail -8 gpurun_out/r2_run14_emu.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
MAAI_PEER_TIMEOUT_S=20 timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n2_flags.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n2_flags.log
grep "DIST_\|rc=\|convergence" gpurun_out/r2_dist_n2_flags.log; grep -i "timeout\|error" gpurun_out/r2_dist_n2_flags.log | head -5
for cfg in "direct:" "staged:MAAI_SYM_DIRECT=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs MAAI_PEER_TIMEOUT_S=20 timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 5 --require-peer --no-secondary > gpurun_out/r2_run14_$name.json 2> gpurun_out/r2_run14_$name.err
  echo "== $name rc=$?"; grep "^\[rank" gpurun_out/r2_run14_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_run14_$name.json')); c=d['config']; print('   ms', round(d['ms_per_step'],4), 'median', round(c['ms_median'],4), 'host', round(c['host_issue_ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'e2e host', round(d['e2e']['pipelined_host_issue_ms_per_step'],4), 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'], d['parity']['loss_rel'])"
done
