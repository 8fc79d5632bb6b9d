mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_group.py -x -q > gpurun_out/s10_group.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/s10_group.log
