"""Timeline of one bush-kernel launch (csrc/hssb_bush.cuh) on the shapes of BASELINE configs 1-2: per bush step
(bushes with the same chain depth) when items were drawn, how long they waited, how long each level took."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import hssb200 as hb
import hss_oracle as o
from test_plan_cpu import to_product_tree
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for name, n, leaf, rmin, rmax, k in (("c1-like", 2001, 64, 9, 20, 16), ("c2-like", 2 ** 16, 64, 13, 20, 64)):
    rng = np.random.default_rng(5)
    cl = o.bisection_cluster(n, leaf)
    h = o.random_hss(cl, cl, rng, rmin, rmax)
    P = hb.pack(to_product_tree(hb, h))
    P.set_option(hb.OPT_FLOW_KERNEL, 0)
    if len(sys.argv) > 1:
        P.set_option(hb.OPT_BUSH_LEVELS, int(sys.argv[1]))
    probe = int(os.environ.get('PROBE', '0'))
    P.debug_bush_trace(probe_item=probe)
    X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    for _ in range(5):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
    torch.cuda.synchronize()
    ops, deps, smem, stages = P.debug_bush_plan(0)
    ncol = (k + 15) // 16
    nb = len(deps)
    tr = P.debug_bush_trace(0, nb * ncol).astype(np.int64)
    depth = []
    for b, ds in enumerate(deps):
        depth.append(1 + max((depth[d] for d in ds), default=0))
    depth = np.repeat(np.array(depth), ncol)
    t0 = tr[:, 1].min()
    print(f"{name}: {nb} bushes x {ncol} column tiles, smem {smem * 8} B, kernel span {(tr[:, 11].max() - t0) / 1e3:.1f} us")
    nlev = np.zeros(nb, dtype=int)
    for op in ops:
        nlev[op.bush] = max(nlev[op.bush], op.level + 1)
    for d in range(1, depth.max() + 1):
        m = depth == d
        r = tr[m]
        lv = int(nlev[np.unique(np.nonzero(m)[0] // ncol)].max())
        ends = r[:, 3:3 + min(lv, 8)]
        prev = np.concatenate([r[:, 2:3], ends[:, :-1]], axis=1)
        print(f"  step {d}: {m.sum()} items, drawn {(r[:, 1].min() - t0) / 1e3:.1f}..{(r[:, 1].max() - t0) / 1e3:.1f} us, deps met {(r[:, 2].min() - t0) / 1e3:.1f}..{(r[:, 2].max() - t0) / 1e3:.1f},"
              f" wait med {np.median(r[:, 2] - r[:, 1]) / 1e3:.1f}, levels med us {np.round(np.median(ends - prev, axis=0) / 1e3, 2).tolist()},"
              f" publish med {np.median(r[:, 11] - ends[:, -1]) / 1e3:.2f}, done {(r[:, 11].min() - t0) / 1e3:.1f}..{(r[:, 11].max() - t0) / 1e3:.1f}")
    pr = P.bush_probe
    print(f"  probe item {probe} (cycles; per level and warp: op start, k loops start, k loops end, first accumulator read, op end, before barrier, after barrier -- relative to level start)")
    for l in range(8):
        if pr[l].any():
            for w in range(8):
                if pr[l, w, 0]:
                    print(f"    level {l} warp {w}:", [int(x - pr[l, w, 0]) if x else None for x in pr[l, w, 1:]])
    P.close()
