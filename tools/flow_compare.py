"""Any-shape trees: one launch per level (graph replay) against the dataflow kernel (HSSB_OPT_FLOW_KERNEL) and the bush
kernel (HSSB_OPT_BUSH_KERNEL) at several cuts of the tree."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import hssb200 as hb
import hss_oracle as o
from test_plan_cpu import to_product_tree
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for name, n, leaf, rmin, rmax, k in (("c1-like n=2001 leaf 64 ranks 9-20", 2001, 64, 9, 20, 16), ("c2-like n=2^16 leaf 64 ranks 13-20", 2 ** 16, 64, 13, 20, 64)):
    rng = np.random.default_rng(5)
    cl = o.bisection_cluster(n, leaf)
    h = o.random_hss(cl, cl, rng, rmin, rmax)
    P = hb.pack(to_product_tree(hb, h))
    P.set_option(hb.OPT_USE_GRAPH, 1)
    X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    fl, by = P.flops(k), P.algorithmic_bytes(k)
    out = []
    variants = [(0, 0, 50), (1, 0, 50), (0, 1, 50), (0, 1, 2 * 16 + 1), (0, 1, 2 * 16 + 2), (0, 1, 3 * 16 + 3), (0, 1, 4 * 16 + 2), (0, 1, 3 * 16 + 1), (0, 1, 50), (1, 0, 50)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        variants = [(1, 0, 50), (0, 1, 50)]
    if len(sys.argv) > 1 and sys.argv[1] == "pdl":   # level launches with / without programmatic dependent launch against the dataflow kernel
        variants = [(1, 0, 1), (0, 0, 1), (0, 0, 9), (1, 0, 1), (0, 0, 1), (0, 0, 9)]
    Yref = None
    for flow, bush, levels in variants:
        P.set_option(hb.OPT_FLOW_KERNEL, flow)
        P.set_option(hb.OPT_BUSH_KERNEL, bush)
        if len(sys.argv) > 1 and sys.argv[1] == "pdl":
            P.set_option(hb.OPT_PDL, levels)
        else:
            P.set_option(hb.OPT_BUSH_LEVELS, levels)
        for _ in range(3):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        l0 = P.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        if Yref is None:
            Yref = Y.clone()
        same = float((Y - Yref).norm() / Yref.norm())
        out.append((("bush %d/%d" % (levels // 16, levels % 16)) if bush else ("flow" if flow else ("levels pdl %d" % P.get_option(hb.OPT_PDL))), (P.launch_count() - l0) // 50, round(ms * 1e3, 1), same))
    t_flop, t_mem = fl / 37.1e12, by / 6.4686e12
    print(f"{name}: flops {fl:.3e} bytes {by:.3e} roofline {max(t_flop, t_mem) * 1e6:.1f} us | (kernel, launches, us, rel. difference to the first):", out)
    P.close()
