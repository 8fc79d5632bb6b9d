"""Do DMMA (FP64 tensor path) and DFMA (FP64 pipe) overlap on B200?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hssb200 as hb
for mode, name in ((0, 'all warps DMMA'), (1, 'all warps DFMA'), (2, 'half DMMA + half DFMA (same iteration count)')):
    print(name, '->', round(hb.measure_peak(7, mode), 2), 'TFLOP/s combined')
