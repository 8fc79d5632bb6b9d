#!/bin/bash
# One gpurun call = several independent checks, each in its own process and under its own timeout
# (a trapped kernel kills only its process).  Everything lands in gpurun_out/<tag>_*.log.
tag=${1:-s}
mkdir -p gpurun_out
echo "== tree tests"; timeout 900 python -m pytest tests/test_gpu_tree.py -x -q > gpurun_out/${tag}_tree.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${tag}_tree.log
echo "== trace"; timeout 300 python tools/tree_trace.py c3 c4 > gpurun_out/${tag}_trace.log 2>&1; echo "rc=$?"; grep total gpurun_out/${tag}_trace.log
echo "== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-extra --variant nosolve > gpurun_out/${tag}_bench_tree.json 2> gpurun_out/${tag}_bench_tree.err; echo "rc=$?"
echo "== bench tree-kernel"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-extra --variant nosolve,tree > gpurun_out/${tag}_bench_treek.json 2> gpurun_out/${tag}_bench_treek.err; echo "rc=$?"
echo "== bench full default"; ( time timeout 900 python bench.py > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err ) 2>&1 | grep real; echo "rc=$?"
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],4), 'launches/prod', d.get('launches_per_product'), 'frac', round(d['product_roofline']['frac_of_roofline'],3), 'parity', d.get('parity_rel_err'))
        print('   ', [(p['name'],p['ms']) for p in d['phases_ms'] if p['ms']>0.02], d.get('tree_ms'))
        if d.get('e2e'): print('    e2e', d['e2e']['ms_per_step'], d['e2e'].get('pageable'))
        for k,v in (d.get('extra') or {}).items(): print('    extra', k, {kk: v.get(kk) for kk in ('ms_per_step','value','parity_rel_err','error')}, (v.get('product_roofline') or {}).get('frac_of_roofline'))
    except Exception as e: print(f, 'ERR', e)
P
