mkdir -p gpurun_out
cp -r hssmatrices.jl_b200/lib /tmp/lib_keep
( cd hssmatrices.jl_b200/csrc && touch *.cu && make -s -j16 DEBUG_MODES=1 2>&1 | grep -v "^ptxas\|^$" | head )
timeout 300 python tools/leaf_bounds.py 2>&1 | tee gpurun_out/s4_bounds.log
