mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -8
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -x -q -k "group or two_gpu" > gpurun_out/g2_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/g2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/g2_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/g2_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'parity', d.get('parity_rel_err'), d['config'].get('exchange'))
print('e2e', d['e2e'])
for k,v in (d.get('extra') or {}).items(): print('extra', k, {kk: v.get(kk) for kk in ('ms_per_step','value','parity_rel_err','parity_leaves_checked','error','n_gpus')}, (v.get('product_roofline') or {}).get('frac_of_roofline'))
P
