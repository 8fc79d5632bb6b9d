"""Leaf-kernel times on config 3 against the leading dimension of X / Y (power-of-two stride vs padded)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
P = hb.synthetic(n, ls, r, seed)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
for pad in (0, 16, 32, 128, 1024, 4096 + 16):
    ld = n + pad
    X = torch.randn((k, ld), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    acc = {}
    for it in range(6):
        P.matmul_dev(X.data_ptr(), ld, Y.data_ptr(), ld, k, stream=s.cuda_stream)
        torch.cuda.synchronize()
        if it:
            for ph in P.phase_times():
                acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 5
    print("ld = n + %5d: leaf_up %.4f ms  leaf_down %.4f ms  total %.4f" % (pad, acc["leaf_up"], acc["leaf_down"], sum(acc.values())))
