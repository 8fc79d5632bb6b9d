mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or random_trees or rectangular or alpha_beta" > gpurun_out/s16_flow_test.log 2>&1; echo "flow tests rc=$?"; tail -4 gpurun_out/s16_flow_test.log
python tools/flow_compare.py 2>&1 | tee gpurun_out/s16_flow.log
