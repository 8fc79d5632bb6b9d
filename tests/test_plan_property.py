"""Property test of the native packer + level scheduler on arbitrary trees: random (unbalanced)
shapes, rectangular nodes, independent row/column ranks including 0, random alpha/beta — the plan
executed by the numpy interpreter must equal the dense expansion (full(), hssmatrix.jl:270-305)
applied to X, for the forward and the transposed product."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import plan_interp
from test_plan_cpu import relerr, to_product_tree

TOL = 1e-12


def random_tree(oracle, rng, depth, root=True):
    """Random HSS tree: leaves at random depths, m != n per leaf, ranks in 0..4."""
    def rk():
        return int(rng.integers(0, 5))

    if depth == 0 or (not root and rng.random() < 0.3):
        m, n = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        D = rng.standard_normal((m, n))
        if root:
            return oracle.hss_leaf(D, rootnode=True)
        return oracle.hss_leaf(D, rng.standard_normal((m, rk())), rng.standard_normal((n, rk())))
    A11 = random_tree(oracle, rng, depth - 1, False)
    A22 = random_tree(oracle, rng, depth - 1, False)
    (kr1, kw1), (kr2, kw2) = oracle.gensize(A11), oracle.gensize(A22)
    B12, B21 = rng.standard_normal((kr1, kw2)), rng.standard_normal((kr2, kw1))
    if root:
        return oracle.hss_branch(A11, A22, B12, B21, rootnode=True)
    kr, kw = rk(), rk()
    return oracle.hss_branch(A11, A22, B12, B21, rng.standard_normal((kr1, kr)), rng.standard_normal((kw1, kw)),
                             rng.standard_normal((kr2, kr)), rng.standard_normal((kw2, kw)))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), depth=st.integers(0, 4), k=st.integers(1, 5),
       alpha=st.sampled_from([1.0, -0.5, 2.25]), beta=st.sampled_from([0.0, 1.0, -1.5]))
def test_plan_equals_dense(hb, oracle, seed, depth, k, alpha, beta):
    rng = np.random.default_rng(seed)
    h = random_tree(oracle, rng, depth)
    assert oracle.checkdims(h)
    m, n = oracle.size(h)
    A = oracle.full(h)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    assert P.shape == (m, n)
    for trans in (False, True):
        op = A.T if trans else A
        X = rng.standard_normal((op.shape[1], k))
        C0 = rng.standard_normal((op.shape[0], k))
        Y = np.asfortranarray(C0.copy()) if beta != 0.0 else np.full(C0.shape, np.nan, order="F")
        plan_interp.run_plan(P, X, Y, alpha, beta, trans=trans)
        ref = alpha * (op @ X) + (beta * C0 if beta != 0.0 else 0.0)
        assert relerr(Y, ref) <= TOL or np.linalg.norm(Y - ref) <= 1e-13
    P.close()
