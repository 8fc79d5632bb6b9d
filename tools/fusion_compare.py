"""North star (2) measured: the "X once" fused leaf variant (HSSB_OPT_LEAF_FUSION) against the two-pass default,
per phase (profiled pass) and per product (graph replay), on the config 3 / 4 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
cfgs = {"c3": (2 ** 20, 128, 32, 64), "c4": (2 ** 20, 128, 64, 128)}
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for name in sys.argv[1:] or ["c3", "c4"]:
    n, ls, r, k = cfgs[name]
    P = hb.synthetic(n, ls, r, 3)
    X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    ref = None
    for fus in (0, 1):
        P.set_option(hb.OPT_LEAF_FUSION, fus)
        P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
        acc = {}
        for it in range(6):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream); torch.cuda.synchronize()
            if it:
                for ph in P.phase_times():
                    acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 5
        P.set_option(hb.OPT_PROFILE, 0); P.set_option(hb.OPT_USE_GRAPH, 1)
        for _ in range(3):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        err = None
        if ref is None: ref = Y.clone()
        else: err = float(torch.linalg.norm(Y - ref) / torch.linalg.norm(ref))
        print(f"{name} fusion={fus}: leaf_up {acc['leaf_up']:.4f} ms  leaf_down {acc['leaf_down']:.4f} ms  leaf sum {acc['leaf_up'] + acc['leaf_down']:.4f} ms  "
              f"product (graph) {ms:.4f} ms  rel.diff to two-pass {err}")
    P.close()
