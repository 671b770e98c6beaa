#!/bin/bash
set -x
mkdir -p gpurun_out
{
python tools/ab_variants.py run 32768 128 30
python tools/ab_variants.py run 8192 128 50
} > gpurun_out/r2_run18_ab.log 2>&1
grep "B=" gpurun_out/r2_run18_ab.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_run18_bench20.json 2> gpurun_out/r2_run18_bench20.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run18_bench20.json')); c=d['config']; e=d['e2e']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks'].get('sampled'))
print('spans', {k:round(v,4) for k,v in c['span_ms_mean'].items() if v}, 'median', c['ms_median'], 'max', c['ms_max'])
print('roofline', d['roofline']['frac'], 'e2e', round(e['ms_per_step'],4), e['schedule'], 'compute span', e['pipelined_compute_span_ms'])
print('run_lengths', {k:(round(v['ms_per_step'],4), v['clocks']['sm_mhz'], v['clocks']['reasons']) for k,v in c['run_lengths'].items()})
PY
