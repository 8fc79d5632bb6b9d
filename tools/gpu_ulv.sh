mkdir -p gpurun_out
HSSB_TEST_ULV_FAST=1 timeout 600 python -m pytest tests/test_gpu_zcaller.py -x -q -k "ulv_fast" > gpurun_out/ulv_fast_test.log 2>&1; echo "fast test rc=$?"; tail -15 gpurun_out/ulv_fast_test.log
timeout 600 python tools/ulv_bench.py 2>&1 | tee gpurun_out/ulv_default.log | head -12
timeout 600 python tools/ulv_bench.py --fast 2>&1 | tee gpurun_out/ulv_fast.log | head -12
