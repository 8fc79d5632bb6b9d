"""DMMA throughput vs warps per SM (one CTA per SM, 16 independent accumulator tiles per warp)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hssb200 as hb
for th in (1, 2, 3, 4, 6, 8):
    print(th, 'warps per scheduler ->', round(hb.measure_peak(3, th), 2), 'TFLOP/s')
print('copy GB/s', hb.measure_peak(2, 1 << 30))
