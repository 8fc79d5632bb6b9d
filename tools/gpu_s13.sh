mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "leaf_fusion" > gpurun_out/s13_fusion_test.log 2>&1; echo "fusion test rc=$?"; tail -5 gpurun_out/s13_fusion_test.log
timeout 300 python tools/fusion_compare.py 2>&1 | tee gpurun_out/s13_fusion.log
timeout 300 python -m pytest tests/test_gpu_ulv.py tests/test_gpu_zcaller.py -x -q > gpurun_out/s13_ulv_tests.log 2>&1; echo "ulv tests rc=$?"; tail -3 gpurun_out/s13_ulv_tests.log
timeout 800 python tools/cauchy_configs.py 2>&1 | tee gpurun_out/s13_cauchy.log
