"""Small products through every kernel family (for compute-sanitizer memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import hssb200 as hb
import hss_oracle as o
from test_plan_cpu import to_product_tree

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
rng = np.random.default_rng(0)
# fixed-shape kernels (second generation), the first generation, the X-once variant, the tree kernel
for (n, ls, r, k) in ((1024, 128, 32, 64), (1024, 128, 64, 20), (1024, 256, 32, 9)):
    h = o.synthetic_hss(n, ls, r, 7)
    X = o.synth_x(7, n, k)
    ref = o.matmul(h, X)
    with hb.synthetic(n, ls, r, 7) as P:
        P.set_option(hb.OPT_PIPELINE_COLS, 1 << 20)
        print("leaf2", (n, ls, r, k), rel(P @ X, ref))
        P.set_option(hb.OPT_LEAF_KERNEL, 1); print("  gen1", rel(P @ X, ref)); P.set_option(hb.OPT_LEAF_KERNEL, 2)
        P.set_option(hb.OPT_LEAF_FUSION, 1); print("  x-once", rel(P @ X, ref)); P.set_option(hb.OPT_LEAF_FUSION, 0)
        P.set_option(hb.OPT_TREE_KERNEL, 1); print("  tree kernel", rel(P @ X, ref)); P.set_option(hb.OPT_TREE_KERNEL, 0)
        print("  A'X twin", rel(P.tmatmul(X), o.matmul(o.adjoint(h), X)))
        Z = P.solve(X); print("  solve (fast form) residual", rel(P @ Z, X))
# any-shape: dataflow kernel and level launches
cl = o.bisection_cluster(777, 50)
h = o.random_hss(cl, cl, rng, 1, 9)
X = rng.standard_normal((777, 5))
P = hb.pack(to_product_tree(hb, h))
print("flow", rel(P @ X, o.matmul(h, X)), "A'X", rel(P.tmatmul(X), o.matmul(o.adjoint(h), X)))
P.set_option(hb.OPT_FLOW_KERNEL, 0)
print("levels", rel(P @ X, o.matmul(h, X)))
P.set_option(hb.OPT_BUSH_KERNEL, 1)   # merge / translate levels as one launch over bushes of the tree
for lv in (2 * 16 + 1, 3 * 16 + 2):
    P.set_option(hb.OPT_BUSH_LEVELS, lv)
    print("bush", lv, rel(P @ X, o.matmul(h, X)), "A'X", rel(P.tmatmul(X), o.matmul(o.adjoint(h), X)), "in use:", P.get_option(hb.OPT_BUSH_KERNEL) == 3)
P.close()
cl = o.bisection_cluster(4096, 64)
h = o.random_hss(cl, cl, rng, 13, 40)    # config-2 shape, ranks up to 40: some blocks exceed the shared-memory budget
X = rng.standard_normal((4096, 33))
P = hb.pack(to_product_tree(hb, h))
P.set_option(hb.OPT_BUSH_KERNEL, 1)
P.set_option(hb.OPT_BUSH_LEVELS, 3 * 16 + 2)
print("bush 4096/64 ranks 13-40", rel(P @ X, o.matmul(h, X)))
P.close()
