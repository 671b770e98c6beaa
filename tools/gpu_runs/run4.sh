#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run4_pytest.log
tail -6 gpurun_out/r2_run4_pytest.log
{
python tools/small_batch_time.py 256 4096
MAAI_FAST_EXT=0 python tools/small_batch_time.py 256 4096
python tools/host_overhead.py
MAAI_FAST_EXT=0 python tools/host_overhead.py
} > gpurun_out/r2_run4_timing.log 2>&1
grep -v Warning gpurun_out/r2_run4_timing.log | grep "pairs=\|b="
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run4_bench.json'))
c=d['config']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['schedule'], 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'])
print('run_lengths', {k:(v['ms_per_step'], v['clocks']['sm_mhz']) for k,v in c['run_lengths'].items()})
print('4096', {k:v for k,v in c.get('configs1_4096_pairs').items() if 'ms' in k or 'torch' in k or 'speed' in k})
print('256', {k:v for k,v in c.get('configs0_256_pairs').items() if 'ms' in k})
PY
