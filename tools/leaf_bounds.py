"""Upper bounds for the leaf kernels on config 3: compute-only (no data waits) and move-only (no DMMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
P = hb.synthetic(n, ls, r, seed)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
for mode, name in ((0, "normal"), (1, "compute-only (no data waits)"), (5, "compute-only, no stores"), (2, "move-only (no DMMA)")):
    P.set_option(hb.OPT_DEBUG, mode)
    acc = {}
    for it in range(6):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        torch.cuda.synchronize()
        if it:
            for ph in P.phase_times():
                acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 5
    print(name, "leaf_up %.4f ms  leaf_down %.4f ms  total %.4f" % (acc["leaf_up"], acc["leaf_down"], sum(acc.values())))
