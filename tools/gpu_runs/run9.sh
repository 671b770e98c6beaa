#!/bin/bash
# 8 GPUs: real-rank parity, bench in both sync modes, configs[4] sweep
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR tests/_dist_gpu_worker.py > gpurun_out/r2_dist_n8_flags.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_n8_flags.log
grep "DIST_\|rc=\|modes" gpurun_out/r2_dist_n8_flags.log
B="bench.py --gpus 8 --steps 40 --warmup 5 --require-peer --no-secondary"
for cfg in "flags:" "barriers:MAAI_PEER_FLAGS=0" "flags_sym:MAAI_FWD_SYM_MULTI=1" "flags_nopdl:MAAI_PDL=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 $TR $B > gpurun_out/r2_bench_n8_$name.json 2> gpurun_out/r2_bench_n8_$name.err
  echo "== $name rc=$?"; grep "^\[rank [07]\]" gpurun_out/r2_bench_n8_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_n8_$name.json')); c=d['config']
    print('   ms', round(d['ms_per_step'],4), 'median', round(c['ms_median'],4), 'frac', round(c['step_frac_bf16_peak'],3), 'sustained', round(c['run_lengths']['sustained']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['schedule'], 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'], d['parity']['loss_rel'], c['gather_mode'], 'launches', d['gpu_launches'])
except Exception as e: print('   failed', e)
PY
done
timeout 900 $TR tools/sweep.py --out gpurun_out/r2_sweep_n8.json > gpurun_out/r2_sweep_n8.log 2>&1; echo "sweep rc=$?"
grep "^d=" gpurun_out/r2_sweep_n8.log | tail -45
