# parity of the fixed-shape kernels + per-phase times after a kernel change
tag=${1:-s3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tree.py -x -q -k "not config2 and not cauchy" > gpurun_out/${tag}_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/${tag}_parity.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-extra --variant nosolve > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],4), 'frac', round(d['product_roofline']['frac_of_roofline'],4), 'parity', d.get('parity_rel_err'))
print([(p['name'],p['ms']) for p in d['phases_ms'] if p['ms']>0.02])
P
