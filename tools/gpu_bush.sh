#!/bin/bash
# bush kernel: parity tests, then timings against the dataflow kernel and the level launches
tag=${1:-bush}
mode=${2:-}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bush or dataflow" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python tools/flow_compare.py $mode > gpurun_out/${tag}_compare.log 2>&1; echo "compare rc=$?"; cat gpurun_out/${tag}_compare.log | tail -20
timeout 600 python tools/bush_trace.py > gpurun_out/${tag}_trace.log 2>&1; echo "trace rc=$?"; cat gpurun_out/${tag}_trace.log | tail -20
