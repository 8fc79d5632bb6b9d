mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/g${N}_bench.json 2> gpurun_out/g${N}_bench.err; echo "bench N=$N rc=$?"
tail -2 gpurun_out/g${N}_bench.err
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/g${N}_bench.json').read().strip().splitlines()[-1])
    print('N=$N ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'parity', d.get('parity_rel_err'), d['config'].get('exchange'), 'cpus', d['config']['host_numa'])
    e=d['e2e']; print('  e2e pinned', round(e['ms_per_step'],2), 'pageable', round(e['pageable']['ms_per_step'],2))
    for k,v in (d.get('extra') or {}).items(): print('  extra', k, {kk: v.get(kk) for kk in ('ms_per_step','value','parity_rel_err','parity_leaves_checked','error')}, (v.get('product_roofline') or {}).get('frac_of_roofline'))
except Exception as ex: print('ERR', ex)
P
done
timeout 600 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -3
