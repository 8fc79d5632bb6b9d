import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hssb200 as hb
print('DMMA register-only peak', round(hb.measure_peak(1, 20000), 2))
for kind, name in ((4, 'LDS.64 fragments, 256-DMMA bodies'), (6, 'LDS.64 fragments, 64-DMMA chunks + syncwarp')):
    for per_sm in (1,):
        print(name, per_sm, 'CTA/SM ->', round(hb.measure_peak(kind, per_sm), 2), 'TFLOP/s')
