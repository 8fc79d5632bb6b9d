"""Per-phase device time (profiled pass: plain launches + events) of the config-3 matrix for a given nrhs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
ks = [int(a) for a in sys.argv[1:]] or [8]
n, ls, r, seed = 2 ** 20, 128, 32, 3
P = hb.synthetic(n, ls, r, seed)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
for k in ks:
    X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    acc = {}
    for it in range(6):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream); torch.cuda.synchronize()
        if it:
            for ph in P.phase_times():
                acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 5
    print("nrhs", k, {a: round(v * 1e3, 1) for a, v in acc.items()}, "sum us", round(sum(acc.values()) * 1e3, 1))
