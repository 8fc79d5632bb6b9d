"""Programmatic dependent launch (HSSB_OPT_PDL) on the uniform shapes: graph-replayed products with and without it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import hssb200 as hb
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
only = os.environ.get('PDL_ONLY')
for name, n, ls, r, k in (("c3", 2 ** 20, 128, 32, 64), ("c4 quarter", 2 ** 20, 128, 64, 128), ("c5 eighth", 2 ** 21, 256, 64, 32), ("c3 nrhs 20", 2 ** 20, 128, 32, 20), ("small n=2^16", 2 ** 16, 128, 32, 64)):
    if only and only not in name:
        continue
    P = hb.synthetic(n, ls, r, 3)
    P.set_option(hb.OPT_USE_GRAPH, 1)
    X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    out, ref = [], None
    for pdl in [int(x) for x in os.environ.get('PDL_SEQ', '0,1,3,0,1,3').split(',')]:
        P.set_option(hb.OPT_PDL, pdl)
        for _ in range(5):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        e0.record()
        for _ in range(reps):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        if ref is None:
            ref = Y.clone()
        out.append((pdl, round(e0.elapsed_time(e1) / reps * 1e3, 1), bool(torch.equal(Y, ref))))
    print(f"{name}: (pdl, us per product, identical)", out)
    P.close()
