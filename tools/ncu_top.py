"""Summarise an .ncu-rep: headline metrics per kernel and the top stalled SASS instructions."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'launch__registers_per_thread',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']
seen = set()
for r in rows[2:]:
    n = r[hdr.index('Kernel Name')]
    if n in seen: continue
    seen.add(n); print('==', n[:90])
    for w in want:
        if w in hdr: print(f'   {w:70s} {r[hdr.index(w)]:>16s} {rows[1][hdr.index(w)]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; kern.append(cur)
    elif cur is not None: cur['rows'].append(r)
seen = set()
for k in kern:
    if k['name'] in seen: continue
    seen.add(k['name'])
    hdr = k['rows'][0]; data = [r for r in k['rows'][1:] if len(r) == len(hdr)]
    si = hdr.index('# Samples'); so = hdr.index('Source'); ie = hdr.index('Instructions Executed')
    sc = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[si]) for r in data)
    print('=====', k['name'][:70], 'samples', tot, 'instrs', len(data))
    agg = {hdr[i][6:]: sum(int(r[i]) for r in data) for i in sc}
    print('   ', {a: b for a, b in sorted(agg.items(), key=lambda kv: -kv[1]) if b > tot * 0.005})
    # opcode histogram of samples
    from collections import defaultdict
    op = defaultdict(int); ex = defaultdict(int)
    for r in data:
        toks = r[so].split(); o = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
        o = o.split('.')[0]; op[o] += int(r[si]); ex[o] += int(r[ie])
    print('    by opcode:', [(o, f'{100*v/tot:.1f}%', ex[o]) for o, v in sorted(op.items(), key=lambda kv: -kv[1])[:14]])
    for r in sorted(data, key=lambda r: -int(r[si]))[:topn]:
        st = {hdr[i][6:]: int(r[i]) for i in sc if int(r[i]) > 0}
        print(f"{int(r[si]):6d} {100*int(r[si])/tot:5.1f}% ex={r[ie]:>9s} {r[so].strip()[:62]:62s} {dict(sorted(st.items(), key=lambda kv:-kv[1])[:3])}")
