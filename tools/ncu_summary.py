"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`."""
import csv, sys
from collections import Counter
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for k in keys:
    if k in hdr:
        print(k, [r[hdr.index(k)][:44] for r in rows[2:]])
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
        vals = [r[i] for r in rows[2:]]
        try:
            if max(float(v) for v in vals) > 800:
                print('  stall', h.replace('smsp__pcsamp_warps_issue_stalled_', ''), vals)
        except ValueError:
            pass
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
sections, cur, seen = [], None, set()
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'data': []}
        sections.append(cur)
        continue
    if len(r) == len(hdr) and r[isamp] != '# Samples':
        cur['data'].append(r)
for s in sections:
    if s['name'] in seen:
        continue
    seen.add(s['name'])
    data = s['data']
    tot = sum(int(r[isamp]) for r in data)
    print('=====', s['name'][:80], 'samples', tot)
    c, ex = Counter(), Counter()
    for r in data:
        toks = r[ia].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        c[op] += int(r[isamp]); ex[op] += int(r[iex])
    for op, n in c.most_common(8):
        print('   %-22s %6d %5.1f%%  executed %d' % (op, n, 100.0 * n / tot, ex[op]))
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:6]:
        print('      ', r[isamp], r[iex], r[ia][:80])
