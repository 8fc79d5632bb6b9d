"""Leaf-kernel generations side by side on config 3/4/5 shapes: per-phase times (profiled pass) and bit-identity."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
cfgs = {"c3": (2 ** 20, 128, 32, 64), "c4": (2 ** 20, 128, 64, 128), "c5": (2 ** 21, 256, 64, 32), "c3k20": (2 ** 20, 128, 32, 20), "c3k256": (2**19, 128, 32, 256)}
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for name in sys.argv[1:] or ["c3"]:
    n, ls, r, k = cfgs[name]
    P = hb.synthetic(n, ls, r, 3)
    P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
    X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    ref = None
    for gen in (1, 2, 3):
        P.set_option(hb.OPT_LEAF_KERNEL, gen)
        acc = {}
        for it in range(6):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
            torch.cuda.synchronize()
            if it:
                for ph in P.phase_times():
                    acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 5
        same = None
        if ref is None: ref = Y.clone()
        else: same = bool(torch.equal(ref, Y))
        print(name, "leaf generation", gen, "leaf_up %.4f ms  leaf_down %.4f ms  total %.4f  bit-identical to gen 1: %s" % (acc["leaf_up"], acc["leaf_down"], sum(acc.values()), same))
    P.close()
