# the driver's round-end sequence: gpu tests, smoke, default bench
tag=${1:-full}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
( time timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'frac', round(d['product_roofline']['frac_of_roofline'],4), 'parity', d.get('parity_rel_err'), 'launches', d.get('launches_per_product'))
print([(p['name'],p['ms']) for p in d['phases_ms'] if p['ms']>0.02])
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pageable']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
for k,v in (d.get('extra') or {}).items(): print('extra', k, {kk: v.get(kk) for kk in ('ms_per_step','value','parity_rel_err','error')}, (v.get('product_roofline') or {}).get('frac_of_roofline'))
print('solve', d.get('ulv_solve'))
P
