#!/bin/bash
# N GPUs (N = $1): final bench line of the round + chained-view K1 span
set -x
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --require-peer > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err
echo "== bench n$N rc=$?"; grep "^\[rank [0]\]" gpurun_out/r2_bench_n${N}_final.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n${N}_final.json')); c=d['config']; e=d['e2e']
print('   ms', round(d['ms_per_step'],4), 'median', round(c['ms_median'],4), 'max', round(c['ms_max'],4), 'frac', round(c['step_frac_bf16_peak'],3), 'value', d['value'])
print('   run_lengths', {k:(round(v['ms_per_step'],4), v['clocks']['sm_mhz'], v['clocks']['reasons']) for k,v in c['run_lengths'].items()})
print('   e2e', round(e['ms_per_step'],4), e['schedule'], 'copies', round(e['copies_only_ms_per_step'],4), 'compute span', e['pipelined_compute_span_ms'], 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'], d['parity']['loss_rel'], c['gather_mode'], c.get('sym_forward_mode'), c.get('peer_order'), 'launches', d['gpu_launches'])
print('   detached', c['hidden1_detached']['ms_per_step'], 'clocks', d['clocks'])
PY
timeout 600 $TR tools/chain_time.py 32768 > gpurun_out/r2_chain_k1_span_n$N.log 2>&1; grep "^W=" gpurun_out/r2_chain_k1_span_n$N.log
