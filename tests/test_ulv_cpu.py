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


def random_square_tree(oracle, rng, depth, root=True):
    """Random HSS tree with square leaves of 1..9 rows at random depths, ranks 0..4 (so that some nodes
    have rank >= size and are handed up whole, ulvfactor.jl:31-37), diagonally shifted."""
    def rk():
        return int(rng.integers(0, 5))

    if depth == 0 or (not root and rng.random() < 0.3):
        m = int(rng.integers(1, 10))
        D = rng.standard_normal((m, m)) + 6.0 * np.eye(m)
        if root:
            return oracle.hss_leaf(D, rootnode=True)
        return oracle.hss_leaf(D, rng.standard_normal((m, rk())) / 3, rng.standard_normal((m, rk())) / 3)
    A11 = random_square_tree(oracle, rng, depth - 1, False)
    A22 = random_square_tree(oracle, rng, depth - 1, False)
    (kr1, kw1), (kr2, kw2) = oracle.gensize(A11), oracle.gensize(A22)
    B12, B21 = rng.standard_normal((kr1, kw2)), rng.standard_normal((kr2, kw1))
    if root:
        return oracle.hss_branch(A11, A22, B12, B21, rootnode=True)
    kr, kw = rk(), rk()
    return oracle.hss_branch(A11, A22, B12, B21, rng.standard_normal((kr1, kr)) / 2, rng.standard_normal((kw1, kw)) / 2,
                             rng.standard_normal((kr2, kr)) / 2, rng.standard_normal((kw2, kw)) / 2)


def test_solve_plan_property(hb, oracle, ulv_oracle):
    """Arbitrary (unbalanced, variable / zero / full rank) square trees: host-team factorisation + interpreted
    solve plan against the dense solve and the oracle restatement."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(seed=st.integers(0, 2 ** 31 - 1), depth=st.integers(0, 4), k=st.integers(1, 4))
    def run(seed, depth, k):
        rng = np.random.default_rng(seed)
        h = random_square_tree(oracle, rng, depth)
        A = oracle.full(h)
        n = A.shape[0]
        B = rng.standard_normal((n, k))
        cond = np.linalg.cond(A)
        if cond > 1e8:
            return
        P = hb.pack(to_product_tree(hb, h), plan_only=True)
        if P.ulv_info.supported == 0:      # a node would eliminate more rows than it has columns
            P.close()
            return
        Z = solve_by_plan(P, B)
        P.close()
        ref = ulv_oracle.ulvfactsolve(h, B)
        assert np.linalg.norm(A @ Z - B) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Z) * max(1.0, np.sqrt(n))
        assert np.linalg.norm(Z - ref) <= 1e-13 * cond * np.linalg.norm(ref)

    run()


def test_options_and_info_on_plan_only_handles(hb, oracle):
    P = hb.synthetic(1024, 128, 16, 2, plan_only=True)
    ui = P.ulv_info
    assert ui.supported == 1 and ui.factored == 0 and ui.pool_bytes > 0 and ui.flops_per_rhs > P.info.flops_per_rhs
    assert P.get_option(hb.OPT_ADJOINT_TWIN) == 1
    P.set_option(hb.OPT_ADJOINT_TWIN, 0)
    assert P.get_option(hb.OPT_ADJOINT_TWIN) == 0
    P.set_option(hb.OPT_PIPELINE_COLS, 12)
    assert P.get_option(hb.OPT_PIPELINE_COLS) == 12
    with pytest.raises(hb.HssbError):
        P.ulv_factor()                      # no CPU fallback for the product path
    with pytest.raises(hb.HssbError):
        P.debug_ulv_pool()                  # not factorised yet
    assert P.debug_ulv_pool(factor_on_host=True).size == ui.pool_bytes // 8
    assert P.ulv_info.factored == 1


def reference_test_matrix():
    """test/runtests.jl:17-30: one-sided Cauchy-like kernel on -1:0.001:1 (n = 2001)."""
    x = np.linspace(-1, 1, 2001)
    d = x[:, None] - x[None, :]
    return np.where(d > 0, 0.001 / np.where(d > 0, d, 1.0), 2.0)


def test_reference_solver_assertion_restated(hb, oracle, ulv_oracle):
    """The reference's own check of `\\` (test/runtests.jl:66-67): hssA = compress(A) at tol 1e-6, leafsize 50,
    rhs = randn(n, 5), x = hssA \\ rhs, x0 = A \\ rhs, ||x0 - x|| / ||x0|| <= 50 * tol — for the oracle
    restatement and for the library's factorisation + solve plan."""
    A = reference_test_matrix()
    tol, c = 1e-6, 50.0
    h = oracle.hss(A, leafsize=50, atol=tol, rtol=tol)
    rhs = np.random.default_rng(1).standard_normal((2001, 5))
    x0 = np.linalg.solve(A, rhs)
    xr = ulv_oracle.ulvfactsolve(h, rhs)
    assert np.linalg.norm(x0 - xr) / np.linalg.norm(x0) <= c * tol
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    Z = solve_by_plan(P, rhs)
    assert np.linalg.norm(x0 - Z) / np.linalg.norm(x0) <= c * tol
    assert np.linalg.norm(xr - Z) / np.linalg.norm(xr) <= 1e-12        # parity with the oracle on the same generators


def test_solve_plan_rectangular_leaves(hb, oracle, ulv_oracle):
    """Square matrix whose row and column cluster trees split differently (leaves with m != n, e.g. 251 x 250
    next to 250 x 251): the reduced blocks stay rectangular up to the root."""
    def cluster(lo, hi, leaf, bias):
        n = hi - lo
        if n <= leaf:
            return oracle.ClusterTree((lo, hi))
        mid = lo + (n + bias) // 2
        return oracle.ClusterTree((lo, hi), cluster(lo, mid, leaf, 1 - bias), cluster(mid, hi, leaf, bias))

    for n, leaf in ((501, 70), (333, 25)):
        rng = np.random.default_rng(n)
        h = oracle.random_hss(cluster(0, n, leaf, 1), cluster(0, n, leaf, 0), rng, 2, 6)
        A = oracle.full(h)
        B = rng.standard_normal((n, 3))
        P = hb.pack(to_product_tree(hb, h), plan_only=True)
        assert P.ulv_info.supported == 1
        Z = solve_by_plan(P, B)
        ref = ulv_oracle.ulvfactsolve(h, B)
        assert np.linalg.norm(A @ Z - B) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Z)
        assert np.linalg.norm(Z - ref) <= 1e-14 * np.linalg.cond(A) * np.linalg.norm(ref)


def solve_golden_files():
    import os
    gdir = os.path.join(os.path.dirname(__file__), "golden", "solve")
    return [os.path.join(gdir, f) for f in sorted(os.listdir(gdir)) if f.endswith(".npz")]


def test_solve_golden_fixtures(hb, oracle, ulv_oracle):
    """Committed vectors from DENSE solves (tests/golden/make_golden_solve.py): they guard the ULV oracle and the
    library's factorisation + solve plan; tolerance = conditioning of the stored matrix."""
    import make_golden
    files = solve_golden_files()
    assert len(files) >= 4
    for f in files:
        z = np.load(f)
        h = make_golden.tree_from_npz(oracle, z)
        tol = 1e-14 * float(z["cond"]) * np.linalg.norm(z["Z"])
        assert np.linalg.norm(ulv_oracle.ulvfactsolve(h, z["B"]) - z["Z"]) <= tol, f
        P = hb.pack(to_product_tree(hb, h), plan_only=True)
        assert np.linalg.norm(solve_by_plan(P, z["B"]) - z["Z"]) <= tol, f


@pytest.mark.parametrize("n,ls,r", [(2048, 128, 16), (2048, 128, 32), (4096, 256, 32), (1024, 128, 64)])
def test_ulv_fast_form_plan(hb, oracle, ulv_oracle, n, ls, r):
    """HSSB_OPT_ULV_FAST (default on; it qualifies on uniform trees): the solve plan in the shapes / padding of the
    product's blocks.  Same solution as the general form (option 0); the phases a fixed-shape kernel can take are
    tagged (leaf-up as V'-like 2r x m, square 2r x 2r merges, r x r top-down steps, leaf-down as D/U-like); fewer
    flops (zloc is folded into the leaf output operator); switching back and forth restores each plan bit for bit."""
    seed = 4
    h = oracle.synthetic_hss(n, ls, r, seed)
    B = oracle.synth_x(seed, n, 3)
    ref = ulv_oracle.ulvfactsolve(h, B)
    P = hb.synthetic(n, ls, r, seed, plan_only=True)
    assert P.get_option(hb.OPT_ULV_FAST) == 2
    Zf = solve_by_plan(P, B)
    P.set_option(hb.OPT_ULV_FAST, 0)
    assert P.get_option(hb.OPT_ULV_FAST) == 0
    Z0 = solve_by_plan(P, B)
    info0 = (P.ulv_info.flops_per_rhs, P.ulv_info.pool_bytes)
    P.set_option(hb.OPT_ULV_FAST, 1)
    assert P.get_option(hb.OPT_ULV_FAST) == 2 and P.ulv_info.factored == 0
    Z1 = solve_by_plan(P, B)
    assert np.array_equal(Z1, Zf)
    tol = 1e-10 * np.linalg.norm(ref)   # these seeded matrices have cond ~ 1e5..1e6; measured 3e-13..5e-13
    assert np.linalg.norm(Z0 - ref) <= tol and np.linalg.norm(Z1 - ref) <= tol
    assert P.ulv_info.flops_per_rhs <= info0[0]
    _, phases, _ = P.debug_plan()
    fast = {(ph.kind, ph.fast) for ph in phases if ph.transposed == 2}
    assert (4, 4) in fast and (3, 3) in fast                      # leaf-down and the r x r top-down steps
    assert ((0, 1) in fast) == (2 * r <= 64) and ((1, 2) in fast) == (2 * r <= 64)   # V'-like leaf-up / square merges need 2r <= 64
    assert (3, 0) in fast                                          # the root stays on the any-shape kernel
    P.set_option(hb.OPT_ULV_FAST, 0)
    assert P.get_option(hb.OPT_ULV_FAST) == 0
    assert np.array_equal(solve_by_plan(P, B), Z0)


def test_ulv_fast_form_needs_uniform_tree(hb, oracle):
    rng = np.random.default_rng(8)
    cl = oracle.bisection_cluster(300, 40)
    P = hb.pack(to_product_tree(hb, oracle.random_hss(cl, cl, rng, 1, 5)), plan_only=True)
    P.set_option(hb.OPT_ULV_FAST, 1)
    assert P.get_option(hb.OPT_ULV_FAST) == 1        # requested, but the plan stays in the default form
    B = rng.standard_normal((300, 2))
    Z = solve_by_plan(P, B)
    assert np.isfinite(Z).all()


def test_singular_matrix_is_reported(hb, oracle):
    """The reference throws SingularException from `D \\ b` (ulvfactor.jl:83) and from the trsm of :48 when a reduced
    block is singular; the library records the pivots its factorisation divides by and returns
    HSSB_ERR_SINGULAR instead of caching factors full of Inf / NaN."""
    rng = np.random.default_rng(3)
    # (1) the root is a leaf and its D is singular (two equal rows)
    D = rng.standard_normal((40, 40))
    D[7] = D[3]
    D[:, 9] = 0.0
    leafroot = hb.HssMatrix.leaf(D, np.zeros((40, 0)), np.zeros((40, 0)), rootnode=True)
    P = hb.pack(leafroot, plan_only=True)
    with pytest.raises(hb.SingularException):
        P.debug_ulv_pool(factor_on_host=True)
    assert P.ulv_info.factored == 0
    # (2) a two-level tree whose reduced root block is exactly zero: all generators zero
    cl = oracle.bisection_cluster(128, 32)
    h = oracle.random_hss(cl, cl, rng, 2, 4)

    def zero(t):
        for f in ("D", "U", "V", "B12", "B21", "R1", "R2", "W1", "W2"):
            a = getattr(t, f, None)
            if isinstance(a, np.ndarray):
                a[...] = 0.0
        if not t.leafnode:
            zero(t.A11)
            zero(t.A22)

    zero(h)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    with pytest.raises(hb.SingularException):
        P.debug_ulv_pool(factor_on_host=True)
    # (3) a regular matrix still factorises
    h = shifted(oracle, oracle.random_hss(cl, cl, rng, 2, 4), 10.0)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    P.debug_ulv_pool(factor_on_host=True)
    assert P.ulv_info.factored == 1
