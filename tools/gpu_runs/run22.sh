#!/bin/bash
# final 1-GPU validation of the round: full GPU suite, smoke, default bench (200 steps), reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_pytest.log
grep -v "Warning\|warn\|^  \|^$" gpurun_out/r2_final_pytest.log | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_final_bench200.json 2> gpurun_out/r2_final_bench200.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench200.json')); c=d['config']; e=d['e2e']
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, d['clocks']['sm_mhz'], d['clocks']['reasons'])
print('roofline', d['roofline']['frac'], d['roofline']['traffic'], 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'])
print('e2e', round(e['ms_per_step'],4), e['schedule'], 'serial', round(e['serial_ms_per_step'],4), 'copies', round(e['copies_only_ms_per_step'],4))
print('run_lengths', {k:(round(v['ms_per_step'],4), v['clocks']['sm_mhz'], v['clocks']['reasons']) for k,v in c['run_lengths'].items()})
print('4096', {k:(round(v,4) if isinstance(v,float) else v) for k,v in c['configs1_4096_pairs'].items() if 'ms_per' in k or 'speed' in k})
print('256', {k:round(v,4) for k,v in c['configs0_256_pairs'].items() if 'ms_per' in k})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ref arm', d['value'], d['cpu_baseline']['kind'], d['ms_per_step'])"
