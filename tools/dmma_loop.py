"""LDS + DMMA inner loop in isolation: simple vs random operand values (data-dependent power?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hssb200 as hb
print('DMMA register-only peak', round(hb.measure_peak(1, 20000), 2))
for arg, name in ((1, 'simple operand values'), (-1, 'random operand values')):
    for kind in (4, 6):
        print('kind', kind, name, '->', round(hb.measure_peak(kind, arg), 2), 'TFLOP/s')
