#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
B="bench.py --gpus 2 --steps 40 --warmup 5 --require-peer --no-secondary --no-parity"
for cfg in "flags:" "flags_bwdbar:MAAI_PEER_FLAGS_BWD=0" "barriers:MAAI_PEER_FLAGS=0" "flags_nosym:MAAI_FWD_SYM_MULTI=0" "barriers_nosym:MAAI_FWD_SYM_MULTI=0 MAAI_PEER_FLAGS=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 $TR $B > gpurun_out/r2_run7_$name.json 2> gpurun_out/r2_run7_$name.err
  echo "== $name rc=$?"; grep "^\[rank" gpurun_out/r2_run7_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_run7_$name.json')); print('   ms', round(d['ms_per_step'],4), 'median', round(d['config']['ms_median'],4), 'sustained', round(d['config']['run_lengths']['sustained']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done
