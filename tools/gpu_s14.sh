mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or random_trees or rectangular or alpha_beta or cauchy or config" > gpurun_out/s14_flow_test.log 2>&1; echo "flow tests rc=$?"; tail -6 gpurun_out/s14_flow_test.log
timeout 800 python tools/cauchy_configs.py 2>&1 | tee gpurun_out/s14_cauchy.log | head -3
