"""Generates the small golden fixtures of the solver, tests/golden/solve/*.npz (run once, committed).

The reference has no golden vectors for `hssA \\ B` and cannot be run here (no Julia), so these come
from the DENSE solve: Z = numpy.linalg.solve(full(hssA), B) — not through the oracle's ULV recursion —
and guard the ULV oracle (oracle/hss_ulv_oracle.py), the library's factorisation + solve plan
(tests/test_ulv_cpu.py) and the CUDA path (tests/test_gpu_zcaller.py).

    python tests/golden/make_golden_solve.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", "..", "oracle"), HERE]


def shift_leaves(h, s):
    if h.leafnode:
        h.D = h.D + s * np.eye(*h.D.shape)
        return
    shift_leaves(h.A11, s)
    shift_leaves(h.A22, s)


def main():
    import hss_oracle as o
    from make_golden import tree_to_dict
    rng = np.random.default_rng(20261018)
    cases = {
        "ragged_n157_l20": dict(n=157, ls=20, k=3, rmin=1, rmax=4),
        "fullrank_n96_l16": dict(n=96, ls=16, k=2, rmin=14, rmax=20),   # ranks around the leaf size: some nodes eliminate nothing
        "zero_rank_n64_l16": dict(n=64, ls=16, k=2, rmin=0, rmax=1),
        "leafroot_n30": dict(n=30, ls=64, k=4, rmin=1, rmax=2),
    }
    os.makedirs(os.path.join(HERE, "solve"), exist_ok=True)
    for name, c in cases.items():
        cl = o.bisection_cluster(c["n"], c["ls"])
        h = o.random_hss(cl, cl, rng, c["rmin"], c["rmax"])
        shift_leaves(h, 4.0 * np.sqrt(c["ls"]))
        A = o.full(h)
        B = rng.standard_normal((c["n"], c["k"]))
        Z = np.linalg.solve(A, B)
        np.savez_compressed(os.path.join(HERE, "solve", name + ".npz"), B=B, Z=Z, cond=np.linalg.cond(A), **tree_to_dict(o, h))
        print(name, "cond %.2e" % np.linalg.cond(A))


if __name__ == "__main__":
    main()
