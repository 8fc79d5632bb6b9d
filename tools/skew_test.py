import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, k = 2 ** 20, 128, 32, 64
P = hb.synthetic(n, ls, r, 3)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
for dbg in (0, 8, 0, 8):
    P.set_option(hb.OPT_DEBUG, dbg)
    acc = {}
    for it in range(8):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        torch.cuda.synchronize()
        if it:
            for ph in P.phase_times():
                acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 7
    print("debug", dbg, "leaf_up %.4f ms  leaf_down %.4f ms  total %.4f" % (acc["leaf_up"], acc["leaf_down"], sum(acc.values())))
