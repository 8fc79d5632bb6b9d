mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ulv.py tests/test_gpu_parity.py -x -q -k "right_division or singular or graph_cache or solver_errors" > gpurun_out/s17.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/s17.log
