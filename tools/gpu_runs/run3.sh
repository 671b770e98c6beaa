#!/bin/bash
set -x
mkdir -p gpurun_out
export MAAI_DEBUG_SEGV=1
timeout 900 python -X faulthandler -m pytest tests/test_gpu_large_batch.py tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2_run3_segv.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run3_segv.log
grep -n "maai: SIGSEGV" -A 30 gpurun_out/r2_run3_segv.log | head -50; tail -3 gpurun_out/r2_run3_segv.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest.log
tail -12 gpurun_out/r2_run3_pytest.log
unset MAAI_DEBUG_SEGV
{
python tools/small_batch_time.py 256 4096
MAAI_PDL=0 python tools/small_batch_time.py 256 4096
python tools/quick_time.py 4096 128 100
MAAI_PDL=0 python tools/quick_time.py 4096 128 100
python tools/quick_time.py 32768 128 30
MAAI_PDL=0 python tools/quick_time.py 32768 128 30
} > gpurun_out/r2_run3_timing.log 2>&1
grep -v Warning gpurun_out/r2_run3_timing.log | grep "lib=\|B="
timeout 900 python bench.py > gpurun_out/r2_run3_bench.json 2> gpurun_out/r2_run3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run3_bench.json'))
c=d['config']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['schedule'], 'parity', d['parity']['ok'], d['parity']['dh_rel_fro'])
print('run_lengths', c['run_lengths'])
print('4096', c.get('configs1_4096_pairs'))
print('256', c.get('configs0_256_pairs'))
print('cpu', d['cpu_baseline'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_b32768_d128.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r2_run3_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ntxent_tile -s 2 -c 2 -o gpurun_out/r2_tile_full -f python tools/prof_step.py 32768 128 2 > gpurun_out/r2_run3_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
