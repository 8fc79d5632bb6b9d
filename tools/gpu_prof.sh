mkdir -p gpurun_out
# every launch of the shipped default schedule (graph replay off so that ncu sees the kernels), device time per launch
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 60 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 5 --no-cpu --no-e2e --no-extra --variant nograph,nosolve > gpurun_out/launches_r02.out 2>&1; echo "launch list rc=$?"
# one product: all kernels with DRAM bytes and pipe utilisation
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"leaf2|node_kernel" -s 54 -c 27 -o gpurun_out/r02_product python tools/phase_breakdown.py 64 > gpurun_out/r02_product_ncu.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/r02_product.ncu-rep --page raw --csv > gpurun_out/r02_product_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_product.ncu-rep --page source --csv > gpurun_out/r02_product_src.csv 2>/dev/null
rm -f gpurun_out/r02_product.ncu-rep
ls -la gpurun_out | tail -8
