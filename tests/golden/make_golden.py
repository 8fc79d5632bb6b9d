"""Generates the small golden fixtures of tests/golden/ (run once, committed).

The reference has no golden vectors for `hssA*X` and cannot be run here (no
Julia), so these vectors come from the dense expansion: Y = alpha*full(hssA)*X
+ beta*C0 computed with numpy on the dense matrix — i.e. NOT through the
oracle's recursion — and are then used to guard the oracle (tests/test_oracle.py)
and the CUDA path (tests/test_gpu_parity.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))


def tree_to_dict(o, h, prefix="t", out=None):
    out = {} if out is None else out
    if h.leafnode:
        out[prefix + ".D"], out[prefix + ".U"], out[prefix + ".V"] = h.D, h.U, h.V
    else:
        for k in ("B12", "B21", "R1", "W1", "R2", "W2"):
            out[f"{prefix}.{k}"] = getattr(h, k)
        tree_to_dict(o, h.A11, prefix + "1", out)
        tree_to_dict(o, h.A22, prefix + "2", out)
    return out


def tree_from_npz(o, z, prefix="t", root=True):
    if prefix + ".D" in z:
        if root:
            return o.hss_leaf(z[prefix + ".D"], rootnode=True)
        return o.hss_leaf(z[prefix + ".D"], z[prefix + ".U"], z[prefix + ".V"])
    A11 = tree_from_npz(o, z, prefix + "1", False)
    A22 = tree_from_npz(o, z, prefix + "2", False)
    if root:
        return o.hss_branch(A11, A22, z[prefix + ".B12"], z[prefix + ".B21"], rootnode=True)
    return o.hss_branch(A11, A22, z[prefix + ".B12"], z[prefix + ".B21"], z[prefix + ".R1"], z[prefix + ".W1"],
                        z[prefix + ".R2"], z[prefix + ".W2"])


def main():
    import hss_oracle as o
    rng = np.random.default_rng(20261017)
    cases = {
        "ragged_n157_l20": dict(n=157, ls=20, k=3, rmin=1, rmax=4, alpha=1.0, beta=0.0),
        "alpha_beta_n96_l16": dict(n=96, ls=16, k=2, rmin=0, rmax=3, alpha=-0.75, beta=1.5),
        "leafroot_n30": dict(n=30, ls=64, k=4, rmin=1, rmax=2, alpha=2.0, beta=0.0),
    }
    for name, c in cases.items():
        cl = o.bisection_cluster(c["n"], c["ls"])
        h = o.random_hss(cl, cl, rng, c["rmin"], c["rmax"])
        X = rng.standard_normal((c["n"], c["k"]))
        C0 = rng.standard_normal((c["n"], c["k"]))
        Y = c["alpha"] * (o.full(h) @ X) + (c["beta"] * C0 if c["beta"] != 0 else 0.0)
        d = tree_to_dict(o, h)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X, C0=C0, Y=Y, alpha=c["alpha"], beta=c["beta"], **d)
        print(name, "ok")


if __name__ == "__main__":
    main()
