"""CPU tests of the native packer + level scheduler (no GPU): the plan the
library would launch is executed by a numpy interpreter and compared with the
oracle restatement of src/matmul.jl:18-62."""
import numpy as np
import pytest

import plan_interp

TOL = 1e-12  # BASELINE.json north_star: relative Frobenius error


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def to_product_tree(hb, h):
    """oracle tree -> product HssMatrix (same field names, distinct class)."""
    if h.leafnode:
        t = hb.HssMatrix.leaf(h.D, h.U, h.V, rootnode=h.rootnode)
        return t
    return hb.HssMatrix.branch(to_product_tree(hb, h.A11), to_product_tree(hb, h.A22), h.B12, h.B21,
                               h.R1, h.W1, h.R2, h.W2, rootnode=h.rootnode)


CASES = [  # (n, leafsize, nrhs, rmin, rmax)
    (2001, 64, 16, 1, 6),   # README shape: 32 leaves of 62/63 rows
    (777, 50, 5, 1, 9),     # odd sizes (48/49-row leaves)
    (1024, 64, 64, 3, 3),
    (100, 200, 3, 1, 3),    # root is a leaf: plain GEMM (matmul.jl:21-22)
    (130, 64, 1, 0, 2),     # ranks may be 0 (hss_blkdiag)
    (5, 1, 2, 1, 2),        # 1x1 leaves
]


@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax", CASES)
def test_plan_matches_oracle(hb, oracle, n, leafsize, nrhs, rmin, rmax):
    rng = np.random.default_rng(n * 7 + leafsize)
    cl = oracle.bisection_cluster(n, leafsize)
    h = oracle.random_hss(cl, cl, rng, rmin, rmax)
    X = rng.standard_normal((n, nrhs))
    ref = oracle.matmul(h, X)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    Y = np.full((n, nrhs), np.nan, order="F")  # beta = 0 must never read Y
    plan_interp.run_plan(P, X, Y)
    assert relerr(Y, ref) <= TOL
    b, f = oracle.algorithmic_counts(h, nrhs)
    assert P.flops(nrhs) == f
    assert P.algorithmic_bytes(nrhs) == b


def test_alpha_beta(hb, oracle):
    rng = np.random.default_rng(5)
    cl = oracle.bisection_cluster(300, 40)
    h = oracle.random_hss(cl, cl, rng)
    X = rng.standard_normal((300, 4))
    C0 = rng.standard_normal((300, 4))
    ref = oracle.mul(C0.copy(), h, X, 0.7, -1.3)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    Y = np.asfortranarray(C0.copy())
    plan_interp.run_plan(P, X, Y, 0.7, -1.3)
    assert relerr(Y, ref) <= TOL


def test_rectangular_and_unbalanced(hb, oracle):
    """m != n per node, different row/col ranks, leaves at different depths
    (prune_leaves!, hssmatrix.jl:325-335) — none of which the reference tests."""
    rng = np.random.default_rng(11)
    rcl = oracle.bisection_cluster(500, 70)
    ccl = oracle.bisection_cluster(333, 47)  # same tree shape (8 leaves), different sizes
    h = oracle.random_hss(rcl, ccl, rng, 1, 7)
    h.A11 = oracle.prune_leaves(h.A11)
    h.sz1 = oracle.size(h.A11)
    h.A22.A11 = oracle.prune_leaves(h.A22.A11)
    h.A22.sz1 = oracle.size(h.A22.A11)
    X = rng.standard_normal((333, 6))
    ref = oracle.full(h) @ X
    assert relerr(oracle.matmul(h, X), ref) <= TOL
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    assert P.shape == (500, 333)
    Y = np.full((500, 6), np.nan, order="F")
    plan_interp.run_plan(P, X, Y)
    assert relerr(Y, ref) <= TOL


BUSH_CASES = [c for c in CASES if c[0] > c[1] * 2] + [(4096, 64, 20, 9, 20), (3000, 40, 33, 0, 12)]   # trees with at least one merge level


@pytest.mark.parametrize("levels", [3 * 16 + 2, 2 * 16 + 1, 1 * 16 + 0, 5 * 16 + 3])
@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax", BUSH_CASES)
def test_bush_plan_matches_oracle(hb, oracle, n, leafsize, nrhs, rmin, rmax, levels):
    """The bush plan of small any-shape trees (csrc/hssb_bush.cuh: a few levels of the recursion of
    matmul.jl:32-62 per work item, intermediate blocks in shared memory), executed by the numpy interpreter
    with the kernel's visibility rules, for the product and for the transposed task table."""
    rng = np.random.default_rng(n * 11 + leafsize)
    cl = oracle.bisection_cluster(n, leafsize)
    h = oracle.random_hss(cl, cl, rng, rmin, rmax)
    X = rng.standard_normal((n, nrhs))
    ref = oracle.matmul(h, X)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    P.set_option(hb.OPT_BUSH_LEVELS, levels)
    assert P.get_option(hb.OPT_BUSH_LEVELS) == levels
    C0 = rng.standard_normal((n, nrhs))
    Y = np.asfortranarray(C0.copy())
    _, st = plan_interp.run_bush_plan(P, X, Y, 0.7, -1.3)
    assert relerr(Y, 0.7 * ref - 1.3 * C0) <= TOL
    depth = int(P.info.depth)
    if depth >= 6 and levels == 3 * 16 + 2:
        assert st["chain"] <= 2 * ((depth + 2) // 3) + 1 < 2 * depth   # the point of the exercise
    if levels == 2 * 16 + 1:
        assert st["staged_a"] == st["staged_b"] == st["ops"]           # default cut: nothing on a bush's critical path reads global memory
    Y = np.full((n, nrhs), np.nan, order="F")
    plan_interp.run_bush_plan(P, X, Y, trans=True)
    assert relerr(Y, oracle.matmul(oracle.adjoint(h), X)) <= TOL


def test_bush_plan_rectangular_and_unbalanced(hb, oracle):
    rng = np.random.default_rng(11)
    rcl = oracle.bisection_cluster(2000, 70)
    ccl = oracle.bisection_cluster(1332, 47)  # same tree shape (32 leaves), different sizes
    h = oracle.random_hss(rcl, ccl, rng, 1, 7)
    h.A11 = oracle.prune_leaves(h.A11)
    h.sz1 = oracle.size(h.A11)
    h.A22.A11 = oracle.prune_leaves(h.A22.A11)
    h.A22.sz1 = oracle.size(h.A22.A11)
    X = rng.standard_normal((1332, 6))
    ref = oracle.full(h) @ X
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    Y = np.full((2000, 6), np.nan, order="F")
    plan_interp.run_bush_plan(P, X, Y)
    assert relerr(Y, ref) <= TOL
    Xt = rng.standard_normal((2000, 4))
    Yt = np.full((1332, 4), np.nan, order="F")
    plan_interp.run_bush_plan(P, Xt, Yt, trans=True)
    assert relerr(Yt, oracle.full(h).T @ Xt) <= TOL


@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax", CASES[:5])
def test_transposed_plan(hb, oracle, n, leafsize, nrhs, rmin, rmax):
    """Y = A' X on the same packed generators == product with the reference's copied adjoint
    (hssmatrix.jl:165-171), which is what `*(A, hssB)` (matmul.jl:14) computes."""
    rng = np.random.default_rng(n + 1)
    rcl = oracle.bisection_cluster(n, leafsize)
    ccl = oracle.bisection_cluster(n + 7, leafsize + 3) if oracle.nleaves_cl(rcl) == oracle.nleaves_cl(oracle.bisection_cluster(n + 7, leafsize + 3)) else rcl
    h = oracle.random_hss(rcl, ccl, rng, rmin, rmax)
    m_, n_ = oracle.size(h)
    X = rng.standard_normal((m_, nrhs))
    ref = oracle.matmul(oracle.adjoint(h), X)
    assert relerr(ref, oracle.full(h).T @ X) <= TOL
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    Y = np.full((n_, nrhs), np.nan, order="F")
    plan_interp.run_plan(P, X, Y, trans=True)
    assert relerr(Y, ref) <= TOL
    C0 = rng.standard_normal((n_, nrhs))
    Y = np.asfortranarray(C0.copy())
    plan_interp.run_plan(P, X, Y, 0.3, 2.0, trans=True)
    assert relerr(Y, 0.3 * ref + 2.0 * C0) <= TOL


@pytest.mark.parametrize("ls,r", [(128, 16), (256, 32)])
def test_adjoint_twin_pool(hb, oracle, ls, r):
    """Uniform trees: the FORWARD plan over the adjoint twin pool (D', U <-> V, B12 <-> B21', R <-> W,
    hssmatrix.jl:165-180) is A' X, single shard and sharded, so hssb_matmul_t runs the fixed-shape kernels."""
    n, seed, k = 16 * ls, 17, 3
    h = oracle.synthetic_hss(n, ls, r, seed)
    X = oracle.synth_x(seed, n, k)
    ref = oracle.matmul(oracle.adjoint(h), X)
    assert relerr(ref, oracle.full(h).T @ X) <= TOL
    P = hb.synthetic(n, ls, r, seed, plan_only=True)
    twin = P.debug_pool_t()
    Y = np.full((n, k), np.nan, order="F")
    plan_interp.run_plan(P, X, Y, pool=twin)
    assert relerr(Y, ref) <= TOL
    # the twin of the twin is the pool itself
    _, _, pool = P.debug_plan()
    assert not np.array_equal(pool, twin)
    for P_ in (2, 4):
        packs = [hb.synthetic(n, ls, r, seed, shard_rank=g, n_shards=P_, plan_only=True) for g in range(P_)]
        rows = n // P_
        Xs = [X[g * rows:(g + 1) * rows] for g in range(P_)]
        Ys = [np.full((rows, k), np.nan, order="F") for _ in range(P_)]
        plan_interp.run_sharded(packs, Xs, Ys, 0.5, 0.0, pools=[p.debug_pool_t() for p in packs])
        assert relerr(np.vstack(Ys), 0.5 * ref) <= TOL


def test_adjoint_twin_needs_uniform_tree(hb, oracle):
    rng = np.random.default_rng(8)
    cl = oracle.bisection_cluster(300, 40)
    P = hb.pack(to_product_tree(hb, oracle.random_hss(cl, cl, rng, 1, 5)), plan_only=True)
    with pytest.raises(hb.HssbError):
        P.debug_pool_t()
    Q = hb.synthetic(1024, 64, 4, 1, plan_only=True)   # uniform, but not a fixed-shape kernel shape
    with pytest.raises(hb.HssbError):
        Q.debug_pool_t()


def test_subblock_is_rooted(hb, oracle):
    """matmul.jl:24: multiplying a sub-block ignores its own translators."""
    rng = np.random.default_rng(3)
    cl = oracle.bisection_cluster(512, 64)
    h = oracle.random_hss(cl, cl, rng)
    sub = h.A11
    X = rng.standard_normal((256, 3))
    ref = oracle.matmul(sub, X)
    P = hb.pack(to_product_tree(hb, sub), plan_only=True)
    Y = np.zeros((256, 3), order="F")
    plan_interp.run_plan(P, X, Y)
    assert relerr(Y, ref) <= TOL


def test_synthetic_twin_bit_identical(hb, oracle):
    """The library's host twin of the device generator produces the same bits
    as oracle.synthetic_hss, block by block."""
    n, ls, r, seed = 1024, 128, 8, 42
    h = oracle.synthetic_hss(n, ls, r, seed)
    P = hb.synthetic(n, ls, r, seed, plan_only=True)
    assert P.info.uniform == 1 and P.info.n_leaves == 8

    def walk(t, node):
        nd = P.node(node)
        if t.leafnode:
            assert nd.is_leaf
            assert np.array_equal(P.block(node, "D"), t.D)
            assert np.array_equal(P.block(node, "U"), t.U)
            assert np.array_equal(P.block(node, "V"), t.V)
            return
        assert np.array_equal(P.block(node, "B12"), t.B12)
        assert np.array_equal(P.block(node, "B21"), t.B21)
        if t.R1.shape[1]:
            assert np.array_equal(P.block(nd.left, "R"), t.R1)
            assert np.array_equal(P.block(nd.right, "R"), t.R2)
            assert np.array_equal(P.block(nd.left, "W"), t.W1)
            assert np.array_equal(P.block(nd.right, "W"), t.W2)
        walk(t.A11, nd.left)
        walk(t.A22, nd.right)

    walk(h, 0)
    X = oracle.synth_x(seed, n, 4)
    Y = np.zeros((n, 4), order="F")
    plan_interp.run_plan(P, X, Y)
    assert relerr(Y, oracle.matmul(h, X)) <= TOL
    assert (P.algorithmic_bytes(4), P.flops(4)) == oracle.synthetic_counts(n, ls, r, 4)


@pytest.mark.parametrize("P_", [2, 4, 8])
def test_sharded_plan(hb, oracle, P_):
    """Subtree sharding: P plans + one all-gather reproduce the
    single-shard product; every shard only holds its own leaves."""
    n, ls, r, seed, k = 2048, 64, 6, 9, 5
    h = oracle.synthetic_hss(n, ls, r, seed)
    X = oracle.synth_x(seed, n, k)
    ref = oracle.matmul(h, X)
    packs = [hb.synthetic(n, ls, r, seed, shard_rank=g, n_shards=P_, plan_only=True) for g in range(P_)]
    rows = n // P_
    Xs = [X[g * rows:(g + 1) * rows] for g in range(P_)]
    Ys = [np.full((rows, k), np.nan, order="F") for _ in range(P_)]
    for g, p in enumerate(packs):
        assert p.local_shape == (rows, rows) and p.info.local_row0 == g * rows
        assert p.info.n_leaves == (n // ls) // P_
    plan_interp.run_sharded(packs, Xs, Ys)
    assert relerr(np.vstack(Ys), ref) <= TOL


def test_sharded_from_host_tree(hb, oracle):
    """Same through the builder path (hssb_builder_add_remote placeholders),
    with variable ranks."""
    rng = np.random.default_rng(21)
    n, k = 1000, 3
    cl = oracle.bisection_cluster(n, 70)
    h = oracle.random_hss(cl, cl, rng, 2, 7)
    X = rng.standard_normal((n, k))
    ref = oracle.matmul(h, X)
    tree = to_product_tree(hb, h)
    packs = [hb.pack(tree, shard_rank=g, n_shards=4, plan_only=True) for g in range(4)]
    Xs, Ys = [], []
    for p in packs:
        r0, m = p.info.local_col0, p.info.local_n
        Xs.append(X[r0:r0 + m])
        Ys.append(np.full((p.info.local_m, k), np.nan, order="F"))
    plan_interp.run_sharded(packs, Xs, Ys)
    assert relerr(np.vstack(Ys), ref) <= TOL


def test_sharding_a_whole_builder_tree(hb, oracle):
    """hssb_group_finalize registers the WHOLE tree once and finalises it once per shard: the packer then turns
    the other shards' subtrees into placeholders itself.  Same plan, same pool as with hand-made
    hssb_builder_add_remote placeholders."""
    import ctypes as C
    rng = np.random.default_rng(22)
    n, k, P_ = 900, 4, 4
    cl = oracle.bisection_cluster(n, 60)
    h = oracle.random_hss(cl, cl, rng, 1, 8)
    X = rng.standard_normal((n, k))
    tree = to_product_tree(hb, h)
    L = hb.lib()
    b = C.c_void_p()
    hb._check(L.hssb_builder_create(C.byref(b)))
    try:
        root = hb._build(b, tree, 0, 1)     # no placeholders registered
        packs = []
        for g in range(P_):
            hd = C.c_void_p()
            hb._check(L.hssb_plan_only(b, root, g, P_, C.byref(hd)))
            packs.append(hb.PackedHss(hd))
    finally:
        L.hssb_builder_destroy(b)
    manual = [hb.pack(tree, shard_rank=g, n_shards=P_, plan_only=True) for g in range(P_)]
    for a, m in zip(packs, manual):
        (ta, pa, poola), (tm, pm, poolm) = a.debug_plan(), m.debug_plan()
        assert len(ta) == len(tm) and len(pa) == len(pm) and np.array_equal(poola, poolm)
        assert all(bytes(x) == bytes(y) for x, y in zip(ta, tm))
    Xs, Ys = [], []
    for p in packs:
        r0, m = p.info.local_col0, p.info.local_n
        Xs.append(X[r0:r0 + m])
        Ys.append(np.full((p.info.local_m, k), np.nan, order="F"))
    plan_interp.run_sharded(packs, Xs, Ys)
    assert relerr(np.vstack(Ys), oracle.matmul(h, X)) <= TOL


def test_save_load_roundtrip(hb, oracle, tmp_path):
    """Packed format as a file: same blocks, same plan after a round trip."""
    rng = np.random.default_rng(31)
    cl = oracle.bisection_cluster(300, 40)
    h = oracle.random_hss(cl, cl, rng, 0, 6)
    P = hb.pack(to_product_tree(hb, h), plan_only=True)
    f = str(tmp_path / "m.hssb")
    P.save(f)
    Q = hb.load(f, device=-1)
    assert Q.shape == P.shape and Q.info.n_nodes == P.info.n_nodes and Q.info.flops_per_rhs == P.info.flops_per_rhs
    for node in range(P.info.n_nodes):
        for kind in range(7):
            assert np.array_equal(P.block(node, kind), Q.block(node, kind))
    X = rng.standard_normal((300, 3))
    Y = np.zeros((300, 3), order="F")
    plan_interp.run_plan(Q, X, Y)
    assert relerr(Y, oracle.matmul(h, X)) <= TOL
    # the file is untrusted input: a node table that is not a tree (here: a node that is its own child, which would
    # send the planner's walks into an endless loop), absurd sizes or a bad shard header are rejected
    import struct
    raw = open(f, "rb").read()
    hdr = struct.calcsize("<8sIIqqiiiiQq")
    assert raw[:7] == b"HSSB200"

    def damaged(off, fmt, value, name):
        g = str(tmp_path / name)
        with open(g, "wb") as fh:
            fh.write(raw[:off] + struct.pack(fmt, value) + raw[off + struct.calcsize(fmt):])
        with pytest.raises(hb.HssbError):
            hb.load(g, device=-1)

    damaged(hdr, "<q", 0, "selfloop.hssb")                 # root.left = 0: a cycle
    damaged(hdr + 8, "<q", 1, "shared.hssb")               # root.right = root.left: a node referenced twice
    damaged(hdr + 4 * 8, "<q", -5, "negsize.hssb")         # root.m < 0
    damaged(16, "<q", 1 << 50, "hugenodes.hssb")           # n_nodes far beyond the file length
    damaged(32, "<i", 7, "badshard.hssb")                  # shard_rank 7 of 1
    with open(f, "r+b") as fh:      # a damaged file must be rejected, not half-loaded
        fh.truncate(200)
    with pytest.raises(hb.HssbError):
        hb.load(f, device=-1)
    with pytest.raises(hb.HssbError):
        hb.load(str(tmp_path / "missing.hssb"), device=-1)


def test_dimension_mismatch(hb, oracle):
    rng = np.random.default_rng(1)
    cl = oracle.bisection_cluster(64, 16)
    tree = to_product_tree(hb, oracle.random_hss(cl, cl, rng))
    with pytest.raises(hb.DimensionMismatch):  # matmul.jl:19
        hb.mul_(np.zeros((64, 2), order="F"), tree, np.zeros((63, 2)))
    with pytest.raises(hb.DimensionMismatch):  # matmul.jl:20
        hb.mul_(np.zeros((64, 3), order="F"), tree, np.zeros((64, 2)))
    with pytest.raises(hb.DimensionMismatch):  # hssmatrix.jl:58
        hb.HssMatrix.branch(tree.A11, tree.A22, tree.B12, tree.B21, np.zeros((1, 2)), np.zeros((1, 2)),
                            np.zeros((1, 3)), np.zeros((1, 2)))
    # builder-level checks: wrong shard layout
    with pytest.raises(hb.HssbError):
        hb.pack(tree, shard_rank=0, n_shards=64, plan_only=True)


def test_no_cpu_fallback(hb, oracle):
    """A plan-only handle (or a box without a B200) must fail loudly."""
    P = hb.synthetic(256, 64, 4, 1, plan_only=True)
    with pytest.raises(hb.HssbError):
        P @ np.zeros((256, 2))
    if hb.device_count() == 0:
        with pytest.raises(hb.HssbError):
            hb.synthetic(256, 64, 4, 1)
