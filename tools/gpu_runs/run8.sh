#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n2_flags.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n2_flags.log
grep "DIST_\|rc=" gpurun_out/r2_dist_n2_flags.log
B="bench.py --gpus 2 --steps 40 --warmup 5 --require-peer --no-secondary --no-parity"
for cfg in "flags:" "barriers:MAAI_PEER_FLAGS=0" "flags_nosym:MAAI_FWD_SYM_MULTI=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 $TR $B > gpurun_out/r2_run8_$name.json 2> gpurun_out/r2_run8_$name.err
  echo "== $name rc=$?"; grep "^\[rank" gpurun_out/r2_run8_$name.err
done
