"""CPU tests of the ULV solver (hssA \\ B, src/ulvfactor.jl:10-107) without a GPU:
the library's node routine (the same __host__ __device__ code the factorisation kernel runs, here as a
single-thread team) factorises a plan-only handle on the host, and the numpy plan interpreter executes
the solve's task table over the resulting factor pool.  Checked against the oracle restatement of
ulvfactor.jl and against the dense solve."""
import numpy as np
import pytest

import plan_interp
from test_plan_cpu import to_product_tree


def shifted(oracle, h, shift):
    """Add shift*I to the leaf diagonal blocks: keeps the HSS structure, makes the matrix well conditioned."""
    if h.leafnode:
        h.D = h.D + shift * np.eye(*h.D.shape)
        return h
    shifted(oracle, h.A11, shift)
    shifted(oracle, h.A22, shift)
    return h


def solve_by_plan(P, B):
    fpool = P.debug_ulv_pool(factor_on_host=True)
    Z = np.full((P.info.n, B.shape[1]), np.nan, order="F")
    plan_interp.run_plan(P, B, Z, trans=2, pool=fpool)
    return Z


CASES = [  # (n, leafsize, nrhs, rmin, rmax)
    (512, 64, 3, 3, 8),
    (2001, 64, 4, 1, 6),     # README shape: 62/63-row leaves
    (777, 50, 2, 1, 9),
    (130, 64, 1, 0, 2),      # ranks may be 0
    (300, 40, 2, 40, 50),    # rank >= leaf size: nothing can be eliminated at the leaves (ulvfactor.jl:31-37)
    (5, 1, 2, 1, 2),         # 1x1 leaves
    (100, 200, 3, 1, 3),     # the root is a leaf: D \ b (ulvfactor.jl:11-12)
    (1024, 128, 5, 32, 32),
]


@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax", CASES)
def test_solve_plan_matches_oracle(hb, oracle, ulv_oracle, n, leafsize, nrhs, rmin, rmax):
    rng = np.random.default_rng(n + 13 * leafsize)
    cl = oracle.bisection_cluster(n, leafsize)
    h = shifted(oracle, oracle.random_hss(cl, cl, rng, rmin, rmax), 4.0 * np.sqrt(leafsize))
    A = oracle.full(h)
    B = rng.standard_normal((n, nrhs))
    ref = ulv_oracle.ulvfactsolve(h, B)
    cond = np.linalg.cond(A)
    assert np.linalg.norm(A @ ref - B) <= 1e-12 * cond * np.linalg.norm(B)     # the oracle itself
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    assert P.ulv_info.supported == 1
    Z = solve_by_plan(P, B)
    # backward error (scale free) and forward parity with the oracle (conditioning dependent)
    assert np.linalg.norm(A @ Z - B) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Z)
    assert np.linalg.norm(Z - ref) <= 1e-14 * cond * np.linalg.norm(ref) + 1e-300


def test_solve_plan_unbalanced_tree(hb, oracle, ulv_oracle):
    """Leaves at different depths (prune_leaves!, hssmatrix.jl:325-335)."""
    rng = np.random.default_rng(2)
    cl = oracle.bisection_cluster(600, 40)
    h = shifted(oracle, oracle.random_hss(cl, cl, rng, 2, 7), 25.0)
    h.A11 = oracle.prune_leaves(h.A11)
    h.sz1 = oracle.size(h.A11)
    h.A22.A11 = oracle.prune_leaves(h.A22.A11)
    h.A22.sz1 = oracle.size(h.A22.A11)
    A = oracle.full(h)
    B = rng.standard_normal((600, 3))
    ref = ulv_oracle.ulvfactsolve(h, B)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    Z = solve_by_plan(P, B)
    assert np.linalg.norm(Z - ref) <= 1e-12 * np.linalg.norm(ref)
    assert np.linalg.norm(A @ Z - B) <= 1e-12 * np.linalg.norm(B)


def test_solve_plan_synthetic_uniform(hb, oracle, ulv_oracle):
    """Uniform synthetic tree (the padded pool layout of the fixed-shape kernels feeds the factorisation)."""
    n, ls, r, seed = 2048, 128, 16, 4
    h = oracle.synthetic_hss(n, ls, r, seed)
    A = oracle.full(h)
    B = oracle.synth_x(seed, n, 3)
    ref = ulv_oracle.ulvfactsolve(h, B)
    P = hb.synthetic(n, ls, r, seed, plan_only=True)
    Z = solve_by_plan(P, B)
    cond = np.linalg.cond(A)
    assert np.linalg.norm(A @ Z - B) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Z)
    assert np.linalg.norm(Z - ref) <= 1e-14 * cond * np.linalg.norm(ref)


def test_solver_not_applicable(hb, oracle):
    rng = np.random.default_rng(3)
    rcl = oracle.bisection_cluster(500, 70)
    ccl = oracle.bisection_cluster(333, 47)
    P = hb.pack(to_product_tree(hb, oracle.random_hss(rcl, ccl, rng, 1, 7)), plan_only=True)   # not square
    assert P.ulv_info.supported == 0
    with pytest.raises(hb.HssbError):
        P.debug_ulv_pool(factor_on_host=True)
    Q = hb.synthetic(1024, 64, 4, 1, shard_rank=0, n_shards=2, plan_only=True)                  # sharded
    assert Q.ulv_info.supported == 0
    S = hb.synthetic(256, 64, 4, 1, plan_only=True)                                             # no CPU fallback
    with pytest.raises(hb.HssbError):
        S.solve(np.zeros((256, 1)))
