"""Parity tests proper (run on a B200 with `pytest -m gpu`): the CUDA path is
driven through the C ABI (include/hssb200.h) and compared with the oracle on
the same seeded inputs.  Bar (BASELINE.json north_star): relative Frobenius
error <= 1e-12."""
import os

import numpy as np
import pytest

from test_plan_cpu import CASES, to_product_tree

pytestmark = pytest.mark.gpu
TOL = 1e-12


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def gpu(hb):
    if hb.device_count() < 1:
        pytest.fail("no B200 visible: the gpu-marked tests need the real device (no CPU fallback)")
    return hb


@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax", CASES + [(4000, 128, 130, 5, 40), (1500, 256, 33, 17, 17)])
def test_random_trees(gpu, oracle, n, leafsize, nrhs, rmin, rmax):
    rng = np.random.default_rng(n * 7 + leafsize)
    cl = oracle.bisection_cluster(n, leafsize)
    h = oracle.random_hss(cl, cl, rng, rmin, rmax)
    X = rng.standard_normal((n, nrhs))
    ref = oracle.matmul(h, X)
    tree = to_product_tree(gpu, h)
    Y = tree @ X                      # `*(hssA, B)`, matmul.jl:13
    assert Y.shape == ref.shape and relerr(Y, ref) <= TOL
    y = tree @ X[:, 0]                # `*(hssA, x::Vector)`, matmul.jl:15
    assert y.shape == (n,) and relerr(y, ref[:, 0]) <= TOL
    tree._packed.close()


def test_alpha_beta_and_uninitialised_c(gpu, oracle):
    rng = np.random.default_rng(5)
    cl = oracle.bisection_cluster(300, 40)
    h = oracle.random_hss(cl, cl, rng)
    X = rng.standard_normal((300, 4))
    C0 = rng.standard_normal((300, 4))
    tree = to_product_tree(gpu, h)
    got = gpu.mul_(np.asfortranarray(C0.copy()), tree, X, 0.7, -1.3)   # mul!, matmul.jl:18
    assert relerr(got, oracle.mul(C0.copy(), h, X, 0.7, -1.3)) <= TOL
    got = gpu.mul_(np.full((300, 4), np.nan, order="F"), tree, X, 2.0, 0.0)  # beta == 0 never reads C
    assert np.isfinite(got).all() and relerr(got, 2.0 * oracle.matmul(h, X)) <= TOL
    with pytest.raises(gpu.DimensionMismatch):
        gpu.mul_(np.zeros((300, 4), order="F"), tree, X[:299])
    with pytest.raises(gpu.DimensionMismatch):
        tree._packed.mul_(np.zeros((299, 4), order="F"), X)


def test_rectangular_unbalanced_subblock(gpu, oracle):
    rng = np.random.default_rng(11)
    rcl = oracle.bisection_cluster(500, 70)
    ccl = oracle.bisection_cluster(333, 47)
    h = oracle.random_hss(rcl, ccl, rng, 1, 7)
    h.A11 = oracle.prune_leaves(h.A11)
    h.sz1 = oracle.size(h.A11)
    X = rng.standard_normal((333, 6))
    assert relerr(to_product_tree(gpu, h) @ X, oracle.full(h) @ X) <= TOL
    sub = h.A22   # rooted(), matmul.jl:24
    Xs = rng.standard_normal((oracle.size(sub)[1], 2))
    assert relerr(to_product_tree(gpu, sub) @ Xs, oracle.matmul(sub, Xs)) <= TOL


def test_transposed_product(gpu, oracle):
    """`*(A::AbstractMatrix, hssB)` (matmul.jl:14) and hssA' * X without the adjoint copy of
    hssmatrix.jl:165-171: second task table over the same packed generators."""
    rng = np.random.default_rng(14)
    rcl = oracle.bisection_cluster(500, 70)
    ccl = oracle.bisection_cluster(333, 47)
    h = oracle.random_hss(rcl, ccl, rng, 1, 7)
    tree = to_product_tree(gpu, h)
    X = rng.standard_normal((500, 6))
    ref = oracle.matmul(oracle.adjoint(h), X)
    P = tree.repack()
    assert relerr(P.tmatmul(X), ref) <= TOL
    A = rng.standard_normal((4, 500))
    got = A @ tree                                   # matmul.jl:14
    assert got.shape == (4, 333) and relerr(got, A @ oracle.full(h)) <= TOL
    C0 = rng.standard_normal((333, 6))
    got = P.mul_(np.asfortranarray(C0.copy()), X, -0.5, 3.0, trans=True)
    assert relerr(got, -0.5 * ref + 3.0 * C0) <= TOL
    with pytest.raises(gpu.DimensionMismatch):
        P.tmatmul(np.zeros((333, 2)))
    # forward product still right after a transposed one (the two plans share the workspaces)
    Xf = rng.standard_normal((333, 3))
    assert relerr(tree @ Xf, oracle.matmul(h, Xf)) <= TOL
    assert P.get_option(gpu.OPT_ADJOINT_TWIN) == 1    # not a uniform tree: no twin was built
    # uniform synthetic trees (padded TMA-ready layout): the first transposed product builds the adjoint
    # twin pool and A' X runs the forward plan / fixed-shape kernels over it; with the option off it
    # runs the any-shape transposed task table over the primary pool
    for n, ls, r, k, seed in ((2048, 128, 32, 5, 9), (4096, 256, 64, 33, 10), (2048, 128, 16, 64, 11)):
        hs = oracle.synthetic_hss(n, ls, r, seed)
        Xs = rng.standard_normal((n, k))
        ref_t, ref_f = oracle.matmul(oracle.adjoint(hs), Xs), oracle.matmul(hs, Xs)
        with gpu.synthetic(n, ls, r, seed) as Ps:
            assert relerr(Ps.tmatmul(Xs), ref_t) <= TOL
            assert Ps.get_option(gpu.OPT_ADJOINT_TWIN) == 2
            assert relerr(Ps @ Xs, ref_f) <= TOL
            C1 = rng.standard_normal((n, k))
            got = Ps.mul_(np.asfortranarray(C1.copy()), Xs, 2.0, -1.0, trans=True)
            assert relerr(got, 2.0 * ref_t - C1) <= TOL
            twin = Ps.debug_pool_t()                  # device-built twin == host transposition of the pool
            Ps.set_option(gpu.OPT_ADJOINT_TWIN, 0)
            assert Ps.get_option(gpu.OPT_ADJOINT_TWIN) == 0
            assert relerr(Ps.tmatmul(Xs), ref_t) <= TOL
            assert relerr(Ps @ Xs, ref_f) <= TOL
        with gpu.synthetic(n, ls, r, seed, plan_only=True) as Pp:
            assert np.array_equal(twin, Pp.debug_pool_t())


def test_save_load_device(gpu, oracle, tmp_path):
    """hssb_save from the device, hssb_load back onto the device: identical products."""
    n, ls, r, k, seed = 2048, 128, 32, 16, 21
    X = oracle.synth_x(seed, n, k)
    f = str(tmp_path / "synth.hssb")
    with gpu.synthetic(n, ls, r, seed) as P:
        Y = P @ X
        P.save(f)
    with gpu.load(f) as Q:
        assert Q.info.uniform == 1
        assert np.array_equal(Q @ X, Y)
    assert relerr(Y, oracle.matmul(oracle.synthetic_hss(n, ls, r, seed), X)) <= TOL


def test_fuzz_arbitrary_trees(gpu, oracle):
    """Random unbalanced trees (rectangular leaves, ranks 0..4, leaves at different depths) through
    the host entry, forward and transposed, against the dense expansion."""
    from test_plan_property import random_tree
    for seed in range(24):
        rng = np.random.default_rng(1000 + seed)
        h = random_tree(oracle, rng, depth=seed % 5)
        A = oracle.full(h)
        tree = to_product_tree(gpu, h)
        k = 1 + seed % 4
        X = rng.standard_normal((A.shape[1], k))
        got = tree @ X
        assert relerr(got, A @ X) <= TOL or np.linalg.norm(got - A @ X) <= 1e-13, seed
        Xt = rng.standard_normal((A.shape[0], k))
        got = tree._packed.tmatmul(Xt)
        assert relerr(got, A.T @ Xt) <= TOL or np.linalg.norm(got - A.T @ Xt) <= 1e-13, seed
        tree._packed.close()


def test_golden_fixtures(gpu, oracle):
    import make_golden
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for f in sorted(x for x in os.listdir(gdir) if x.endswith(".npz")):
        z = np.load(os.path.join(gdir, f))
        h = make_golden.tree_from_npz(oracle, z)
        got = gpu.mul_(np.asfortranarray(z["C0"].copy()), to_product_tree(gpu, h), z["X"], float(z["alpha"]), float(z["beta"]))
        assert relerr(got, z["Y"]) <= TOL, f


def test_readme_cauchy_config1(gpu, oracle):
    """BASELINE config 1: README kernel n=2001, hss(A, leafsize=64, atol=rtol=1e-6), nrhs 16."""
    A = oracle.cauchy_matrix(2001)
    h = oracle.hss(A, 64, 1e-6, 1e-6)
    X = np.random.default_rng(2001).standard_normal((2001, 16))
    Y = to_product_tree(gpu, h) @ X
    assert relerr(Y, oracle.matmul(h, X)) <= TOL
    assert relerr(Y, A @ X) <= 50e-6       # the reference's own assertion, runtests.jl:63-64


def test_cauchy_config2(gpu, oracle):
    """BASELINE config 2: Cauchy kernel n = 2^16, leafsize 64, atol = rtol = 1e-9, nrhs 64.  A dense
    2^16 x 2^16 matrix (34 GB) cannot be compressed directly; like hss(::LinearMap) in the reference
    (hssmatrix.jl:86) the fixture goes through the randomized compression, restated in the oracle
    (its sampling products A*Omega run blockwise on the GPU through torch to keep the fixture cheap)."""
    n, k = 2 ** 16, 64
    op = oracle.cauchy_operator(n, device="cuda")
    cl = oracle.bisection_cluster(n, 64)
    h = oracle.randcompress(op, cl, cl, 30, 1e-9, 1e-9, rng=np.random.default_rng(16))
    assert oracle.nleaves(h) == 1024 and oracle.checkdims(h)
    X = np.random.default_rng(64).standard_normal((n, k))
    ref = oracle.matmul(h, X)
    tree = to_product_tree(gpu, h)
    Y = tree @ X
    assert relerr(Y, ref) <= TOL
    assert relerr(Y, op.matmat(X)) <= 1e-6      # quality of the compressed fixture, not of the product
    info = tree._packed.info
    assert info.n_leaves == 1024 and info.uniform == 0   # variable ranks: the any-shape kernel
    tree._packed.close()


@pytest.mark.parametrize("n,ls,r,k", [(4096, 128, 32, 64), (8192, 128, 64, 128), (4096, 256, 64, 32), (2048, 128, 32, 7),
                                      (1000, 128, 32, 64), (4096, 64, 16, 20)])
def test_synthetic_device_generated(gpu, oracle, n, ls, r, k):
    """Device generator is bit-identical to the host twin; fixed-shape and
    generic kernels both match the oracle."""
    seed = 1234 + n
    h = oracle.synthetic_hss(n, ls, r, seed)
    X = oracle.synth_x(seed, n, k)
    ref = oracle.matmul(h, X)
    with gpu.synthetic(n, ls, r, seed) as P:
        nd = P.node(0)
        assert np.array_equal(P.block(0, "B12"), h.B12)
        assert P.block(nd.left, "W").shape == (r, 0)   # children of the root carry no translators
        if not h.A11.leafnode:
            assert np.array_equal(P.block(P.node(nd.left).left, "W"), h.A11.W1)
            assert np.array_equal(P.block(P.node(nd.left).right, "R"), h.A11.R2)
        leaf = h
        node = 0
        while not leaf.leafnode:
            leaf, node = leaf.A22, P.node(node).right
        assert np.array_equal(P.block(node, "D"), leaf.D) and np.array_equal(P.block(node, "V"), leaf.V)
        Y = P @ X
        assert relerr(Y, ref) <= TOL
        P.set_option(gpu.OPT_FORCE_GENERIC, 1)
        Yg = P @ X
        assert relerr(Yg, ref) <= TOL
        P.set_option(gpu.OPT_FORCE_GENERIC, 0)
        P.set_option(gpu.OPT_USE_GRAPH, 1)
        for _ in range(2):
            assert relerr(P @ X, ref) <= TOL


def test_device_entry_and_synthetic_rhs(gpu, oracle):
    """hssb_matmul_dev on torch-owned device memory with ld > rows, on torch's stream."""
    import torch
    n, ls, r, k, seed = 4096, 128, 32, 64, 99
    ldx, ldy = n + 8, n + 24
    with gpu.synthetic(n, ls, r, seed) as P:
        X = torch.zeros((k, ldx), dtype=torch.float64, device="cuda")   # column-major n x k, ld = ldx
        Y = torch.full((k, ldy), float("nan"), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        st = torch.cuda.current_stream().cuda_stream
        rc = gpu.lib().hssb_synthetic_rhs(seed, n, k, 0, n, X.data_ptr(), ldx, 0, st)
        assert rc == 0
        P.matmul_dev(X.data_ptr(), ldx, Y.data_ptr(), ldy, k, stream=st)
        torch.cuda.synchronize()
        Xh = X.cpu().numpy()[:, :n].T
        assert np.array_equal(Xh, oracle.synth_x(seed, n, k))
        ref = oracle.matmul(oracle.synthetic_hss(n, ls, r, seed), Xh)
        Yh = Y.cpu().numpy()
        assert relerr(Yh[:, :n].T, ref) <= TOL
        assert np.isnan(Yh[:, n:]).all()      # padding rows untouched
        assert P.launch_count() > 0


@pytest.mark.parametrize("k", [1, 9, 63, 65, 200])
def test_uniform_tree_any_nrhs(gpu, oracle, k):
    """Fixed-shape kernels with partial / multiple column tiles (nrhs not a multiple of anything)."""
    n, ls, r, seed = 2048, 128, 32, 77
    ref = oracle.matmul(oracle.synthetic_hss(n, ls, r, seed), oracle.synth_x(seed, n, k))
    with gpu.synthetic(n, ls, r, seed) as P:
        assert all(p["fast"] for p in P.phase_times())
        assert relerr(P @ oracle.synth_x(seed, n, k), ref) <= TOL


def test_unaligned_x_falls_back_to_generic(gpu, oracle):
    """X with an odd leading dimension or an 8-byte-only aligned base cannot be read by TMA /
    16-byte copies: the leaf phases must take the any-shape kernel and still be exact."""
    import torch
    n, ls, r, k, seed = 2048, 128, 32, 16, 5
    ref = oracle.matmul(oracle.synthetic_hss(n, ls, r, seed), oracle.synth_x(seed, n, k))
    Xh = torch.from_numpy(np.ascontiguousarray(oracle.synth_x(seed, n, k).T))
    with gpu.synthetic(n, ls, r, seed) as P:
        for ldx, shift in ((n + 1, 0), (n + 2, 1)):
            buf = torch.zeros(k * ldx + 8, dtype=torch.float64, device="cuda")
            X = buf[shift:shift + k * ldx].view(k, ldx)
            X[:, :n] = Xh.cuda()
            Y = torch.full((k, n), float("nan"), dtype=torch.float64, device="cuda")
            torch.cuda.synchronize()
            P.matmul_dev(X.data_ptr(), ldx, Y.data_ptr(), n, k, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert relerr(Y.cpu().numpy().T, ref) <= TOL


def test_full_size_config3_properties(gpu, oracle):
    """BASELINE config 3 at full size (n=2^20, leaf 128, rank 32, nrhs 64):
    (a) sparse-support right-hand side checked leaf by leaf against the lazily
    evaluated oracle, (b) linearity A(aX1 + bX2) = aAX1 + bAX2 on dense X."""
    import torch
    n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
    with gpu.synthetic(n, ls, r, seed) as P:
        assert P.info.uniform == 1 and P.info.n_leaves == 8192 and P.info.depth == 13
        assert (P.algorithmic_bytes(k), P.flops(k)) == oracle.synthetic_counts(n, ls, r, k)
        st = torch.cuda.current_stream().cuda_stream
        # (a)
        s_lo, s_len = 5 * ls + 17, 3 * ls
        Xs = np.random.default_rng(0).standard_normal((s_len, k))
        X = torch.zeros((k, n), dtype=torch.float64, device="cuda")
        X[:, s_lo:s_lo + s_len] = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
        Y = torch.empty((k, n), dtype=torch.float64, device="cuda")
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
        torch.cuda.synchronize()
        targets = [0, 5 * ls, 6 * ls, 7 * ls, 8 * ls, 4096 * ls, 8191 * ls, 5000 * ls]
        lazy = oracle.LazySyntheticHss(n, ls, r, seed).rows(targets, s_lo, Xs)
        for lo, yref in lazy.items():
            got = Y[:, lo:lo + ls].cpu().numpy().T
            assert relerr(got, yref) <= TOL, lo
        # (b)
        g = torch.Generator(device="cuda").manual_seed(1)
        X1 = torch.randn((k, n), dtype=torch.float64, device="cuda", generator=g)
        X2 = torch.randn((k, n), dtype=torch.float64, device="cuda", generator=g)
        Y1, Y2, Y3 = (torch.empty((k, n), dtype=torch.float64, device="cuda") for _ in range(3))
        P.matmul_dev(X1.data_ptr(), n, Y1.data_ptr(), n, k, stream=st)
        P.matmul_dev(X2.data_ptr(), n, Y2.data_ptr(), n, k, stream=st)
        X3 = 0.5 * X1 - 2.0 * X2
        P.matmul_dev(X3.data_ptr(), n, Y3.data_ptr(), n, k, stream=st)
        torch.cuda.synchronize()
        lin = 0.5 * Y1 - 2.0 * Y2
        assert (torch.linalg.norm(Y3 - lin) / torch.linalg.norm(lin)).item() <= TOL
        # (c) adjoint identity <X2, A X1> = <A' X2, X1> (A' through the adjoint twin pool), and the twin
        # against the any-shape transposed task table over the primary pool
        P.matmul_dev(X2.data_ptr(), n, Y2.data_ptr(), n, k, stream=st, trans=True)
        torch.cuda.synchronize()
        assert P.get_option(gpu.OPT_ADJOINT_TWIN) == 2
        lhs, rhs = (X2 * Y1).sum().item(), (Y2 * X1).sum().item()
        assert abs(lhs - rhs) <= TOL * (torch.linalg.norm(X2) * torch.linalg.norm(Y1)).item()
        P.set_option(gpu.OPT_ADJOINT_TWIN, 0)
        P.matmul_dev(X2.data_ptr(), n, Y3.data_ptr(), n, k, stream=st, trans=True)
        torch.cuda.synchronize()
        assert (torch.linalg.norm(Y3 - Y2) / torch.linalg.norm(Y2)).item() <= TOL
        # fixed-shape kernels vs the generic kernel on the same dense input
        P.set_option(gpu.OPT_FORCE_GENERIC, 1)
        P.matmul_dev(X1.data_ptr(), n, Y2.data_ptr(), n, k, stream=st)
        torch.cuda.synchronize()
        assert (torch.linalg.norm(Y2 - Y1) / torch.linalg.norm(Y1)).item() <= TOL


@pytest.mark.parametrize("name,n,ls,r,k", [("config 4", 2 ** 22, 128, 64, 128), ("config 5", 2 ** 24, 256, 64, 32)])
def test_full_size_configs_4_and_5(gpu, oracle, name, n, ls, r, k):
    """BASELINE configs 4 and 5 at FULL size on one GPU (15 GB / 64 GB of generators, generated on the device):
    sparse-support right-hand side checked leaf by leaf against the lazily evaluated oracle (leaves next to the
    support, in the other half of the tree and at both ends), and linearity on dense inputs."""
    import torch
    seed = 3
    free, _ = torch.cuda.mem_get_info()
    need = 8 * (n // ls) * (ls * ls + 2 * ls * r) * 1.1 + 5 * 8 * n * k
    if free < need:
        pytest.skip(f"{name} needs {need / 1e9:.0f} GB of device memory, {free / 1e9:.0f} GB free")
    with gpu.synthetic(n, ls, r, seed) as P:
        L = n // ls
        assert P.info.uniform == 1 and P.info.n_leaves == L
        assert (P.algorithmic_bytes(k), P.flops(k)) == oracle.synthetic_counts(n, ls, r, k)
        st = torch.cuda.current_stream().cuda_stream
        s_lo, s_len = 5 * ls + 17, 3 * ls
        Xs = np.random.default_rng(0).standard_normal((s_len, k))
        X = torch.zeros((k, n), dtype=torch.float64, device="cuda")
        X[:, s_lo:s_lo + s_len] = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
        Y = torch.empty((k, n), dtype=torch.float64, device="cuda")
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
        torch.cuda.synchronize()
        targets = [0, 5 * ls, 6 * ls, 7 * ls, 8 * ls, (L // 2) * ls, (L - 1) * ls, (5 * L // 8 + 3) * ls]
        lazy = oracle.LazySyntheticHss(n, ls, r, seed).rows(targets, s_lo, Xs)
        for lo, yref in lazy.items():
            got = Y[:, lo:lo + ls].cpu().numpy().T
            assert relerr(got, yref) <= TOL, (name, lo)
        g = torch.Generator(device="cuda").manual_seed(1)
        X1 = torch.randn((k, n), dtype=torch.float64, device="cuda", generator=g)
        X2 = torch.randn((k, n), dtype=torch.float64, device="cuda", generator=g)
        Y1, Y2 = (torch.empty((k, n), dtype=torch.float64, device="cuda") for _ in range(2))
        P.matmul_dev(X1.data_ptr(), n, Y1.data_ptr(), n, k, stream=st)
        P.matmul_dev(X2.data_ptr(), n, Y2.data_ptr(), n, k, stream=st)
        X1.mul_(0.5).add_(X2, alpha=-2.0)
        P.matmul_dev(X1.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
        torch.cuda.synchronize()
        Y1.mul_(0.5).add_(Y2, alpha=-2.0)
        assert (torch.linalg.norm(Y - Y1) / torch.linalg.norm(Y1)).item() <= TOL
    torch.cuda.empty_cache()


def test_two_gpu_sharded(gpu, oracle):
    """Subtree sharding over NCCL (skipped on a 1-GPU box; the plan itself is
    covered on CPU by tests/test_plan_cpu.py::test_sharded_plan)."""
    if gpu.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(root, "tests", "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SHARDED_OK" in out.stdout


def test_pageable_host_memory_ring(gpu, oracle):
    """hssb_matmul on pageable caller memory (what Julia's `similar(B, ...)`, matmul.jl:13, hands over) goes
    through the library's pinned slot rings and worker threads (csrc/hssb_hostpipe.h).  HSSB_OPT_HOST_BOUNCE = 2
    forces that path for a small matrix: strided X / Y (leading dimension > rows: column-by-column pieces),
    dense ones (whole-panel pieces), beta != 0 (Y travels both ways) -- identical to the direct path."""
    rng = np.random.default_rng(77)
    n, k = 777, 9
    cl = oracle.bisection_cluster(n, 50)
    h = oracle.random_hss(cl, cl, rng, 1, 9)
    tree = to_product_tree(gpu, h)
    P = tree.repack()
    lib = gpu.lib()
    Xbig = np.asfortranarray(rng.standard_normal((n + 13, k)))
    Ybig = np.asfortranarray(rng.standard_normal((n + 5, k)))
    ref = oracle.mul(Ybig[:n].copy(), h, Xbig[:n], 0.7, -1.3)
    outs = []
    for mode in (0, 2):
        P.set_option(gpu.OPT_HOST_BOUNCE, mode)
        Yc = Ybig.copy(order="F")
        rc = lib.hssb_matmul(P._h, n, n, k, Xbig.ctypes.data, n + 13, Yc.ctypes.data, n + 5, 0.7, -1.3)
        assert rc == 0, lib.hssb_last_error()
        assert P.get_option(gpu.OPT_LAST_BOUNCE) == (3 if mode else 0)
        assert relerr(Yc[:n], ref) <= TOL
        assert np.array_equal(Yc[n:], Ybig[n:])      # rows beyond the matrix are not touched
        outs.append(Yc)
    assert np.array_equal(outs[0], outs[1])
    Xd = np.asfortranarray(Xbig[:n])
    Yd = np.full((n, k), np.nan, order="F")
    P.mul_(Yd, Xd)                                    # dense panels, beta == 0 never reads Y
    assert P.get_option(gpu.OPT_LAST_BOUNCE) == 3 and relerr(Yd, oracle.matmul(h, Xd)) <= TOL
    P.set_option(gpu.OPT_HOST_BOUNCE, 1)
    P.mul_(Yd, Xd)                                    # automatic: a call this small is not staged
    assert P.get_option(gpu.OPT_LAST_BOUNCE) == 0
    P.close()


def test_pageable_large_call_is_staged(gpu, oracle):
    """A config-3-sized call (64 MiB and more) on numpy memory takes the ring by default and matches the
    pinned-memory result bit for bit."""
    import torch
    n, ls, r, k, seed = 2 ** 17, 128, 32, 32, 5
    with gpu.synthetic(n, ls, r, seed) as P:
        X = oracle.synth_x(seed, n, k)
        Yp = np.empty((n, k), order="F")
        P.mul_(Yp, X)
        assert P.get_option(gpu.OPT_LAST_BOUNCE) == 3
        Xh = torch.empty((k, n), dtype=torch.float64).pin_memory()
        Yh = torch.empty((k, n), dtype=torch.float64).pin_memory()
        Xh.numpy().T[:] = X
        P.mul_(Yh.numpy().T, Xh.numpy().T)
        assert P.get_option(gpu.OPT_LAST_BOUNCE) == 0
        assert np.array_equal(Yh.numpy().T, Yp)


@pytest.mark.parametrize("r", [32, 64])
def test_leaf_fusion_x_once_variant(gpu, oracle, r):
    """BASELINE north star (2): the fused leaf kernel that reads each X block ONCE for D X and V' X
    (HSSB_OPT_LEAF_FUSION = 1, csrc/hssb_leafx.cuh) against the two-pass default and the oracle, incl. alpha / beta
    (the fused leaf-up parks alpha D X + beta Y in Y, the leaf-down adds alpha U F)."""
    n, ls, seed = 8192, 128, 90 + r
    h = oracle.synthetic_hss(n, ls, r, seed)
    rng = np.random.default_rng(r)
    with gpu.synthetic(n, ls, r, seed) as P:
        for k in (1, 20, 64, 130):
            X = oracle.synth_x(seed + k, n, k)
            ref = oracle.matmul(h, X)
            P.set_option(gpu.OPT_LEAF_FUSION, 0)
            l0 = P.launch_count()
            Y0 = P @ X
            n0 = P.launch_count() - l0
            P.set_option(gpu.OPT_LEAF_FUSION, 1)
            assert P.get_option(gpu.OPT_LEAF_FUSION) == 1
            l0 = P.launch_count()
            Y1 = P @ X
            assert P.launch_count() - l0 == n0            # same number of launches: both leaf kernels are replaced
            assert relerr(Y1, ref) <= TOL and relerr(Y1, Y0) <= 1e-14, (r, k)
            C0 = rng.standard_normal((n, k))
            got = P.mul_(np.asfortranarray(C0.copy()), X, 0.7, -1.3)
            assert relerr(got, oracle.mul(C0.copy(), h, X, 0.7, -1.3)) <= TOL
            got = P.mul_(np.full((n, k), np.nan, order="F"), X, 2.0, 0.0)   # beta == 0 never reads C
            assert np.isfinite(got).all() and relerr(got, 2.0 * ref) <= TOL
        assert relerr(P.tmatmul(X), oracle.matmul(oracle.adjoint(h), X)) <= TOL  # over the adjoint twin pool as well


@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax", [(2001, 64, 16, 9, 20), (777, 50, 5, 1, 9), (4000, 128, 130, 5, 40), (300, 40, 3, 0, 2)])
def test_dataflow_kernel_any_shape_trees(gpu, oracle, n, leafsize, nrhs, rmin, rmax):
    """Trees no fixed-shape kernel applies to run the whole product as ONE persistent dataflow launch
    (HSSB_OPT_FLOW_KERNEL, csrc/hssb_flow.cuh): same tiles, same arithmetic as one launch per level -> identical
    results, one kernel per product; also for A' X on the any-shape transposed task table, with alpha / beta, and
    on repeated calls (the counters are reset by every launch)."""
    rng = np.random.default_rng(n + nrhs)
    rcl = oracle.bisection_cluster(n, leafsize)
    h = oracle.random_hss(rcl, rcl, rng, rmin, rmax)
    X = rng.standard_normal((n, nrhs))
    ref = oracle.matmul(h, X)
    tree = to_product_tree(gpu, h)
    P = tree.repack()
    P.set_option(gpu.OPT_PIPELINE_COLS, 1 << 20)      # one column block per host call: launches are countable
    P.set_option(gpu.OPT_BUSH_KERNEL, 0)              # small trees would go to the bush kernel (next test)
    P.set_option(gpu.OPT_FLOW_KERNEL, 0)
    l0 = P.launch_count()
    Y0 = P @ X
    per_level = P.launch_count() - l0
    T0 = P.tmatmul(X)
    P.set_option(gpu.OPT_FLOW_KERNEL, 1)
    for _ in range(3):
        l0 = P.launch_count()
        Y1 = P @ X
        assert P.launch_count() - l0 == 1 and per_level > 1
        assert np.array_equal(Y1, Y0)
    assert P.get_option(gpu.OPT_FLOW_KERNEL) == 2
    assert relerr(Y1, ref) <= TOL
    l0 = P.launch_count()
    T1 = P.tmatmul(X)
    assert P.launch_count() - l0 == 1
    assert np.array_equal(T1, T0) and relerr(T1, oracle.matmul(oracle.adjoint(h), X)) <= TOL
    C0 = rng.standard_normal((n, nrhs))
    got = P.mul_(np.asfortranarray(C0.copy()), X, 0.7, -1.3)
    assert relerr(got, oracle.mul(C0.copy(), h, X, 0.7, -1.3)) <= TOL
    X2 = rng.standard_normal((n, 2 * nrhs + 1))       # more column tiles than before: the counters grow
    assert relerr(P @ X2, oracle.matmul(h, X2)) <= TOL
    # automatic (the default): the host entry replays a CUDA graph, where one launch per level with programmatic dependent
    # launch is the faster schedule; plain launches on caller-owned device arrays take the dataflow kernel
    P.set_option(gpu.OPT_FLOW_KERNEL, 2)
    l0 = P.launch_count()
    Y2 = P @ X
    assert P.launch_count() - l0 == per_level and np.array_equal(Y2, Y0)
    import torch
    Xd = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    Yd = torch.empty_like(Xd)
    st = torch.cuda.current_stream().cuda_stream
    l0 = P.launch_count()
    P.matmul_dev(Xd.data_ptr(), n, Yd.data_ptr(), n, nrhs, stream=st)
    torch.cuda.synchronize()
    assert P.launch_count() - l0 == 1 and np.array_equal(Yd.cpu().numpy().T, Y0)
    P.close()


@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax,levels", [
    (2001, 64, 16, 9, 20, 50), (777, 50, 5, 1, 9, 50), (4096, 64, 64, 13, 40, 50), (4096, 64, 33, 13, 20, 33), (300, 40, 3, 0, 2, 16),
    (4000, 128, 130, 5, 40, 50), (3000, 40, 17, 0, 12, 83)])
def test_bush_kernel_small_trees(gpu, oracle, n, leafsize, nrhs, rmin, rmax, levels):
    """Small any-shape trees (BASELINE configs 1-2): every merge / translate level between the two leaf launches
    runs as ONE launch over whole bushes of the tree (HSSB_OPT_BUSH_KERNEL, csrc/hssb_bush.cuh): a few levels of
    matmul.jl:39 / :52-56 per CTA out of shared memory, warp-sized tasks
    (on the FP64 FMA pipe: a warp working alone is bound by the latency of a dependent DMMA): agrees with one launch per
    level to rounding, and with itself bit for bit; also A' X on the transposed task table, alpha / beta, repeated calls, other cuts of the tree (HSSB_OPT_BUSH_LEVELS) and, forced (= 2), a tree
    with 128-row leaves."""
    rng = np.random.default_rng(n + nrhs)
    rcl = oracle.bisection_cluster(n, leafsize)
    h = oracle.random_hss(rcl, rcl, rng, rmin, rmax)
    X = rng.standard_normal((n, nrhs))
    ref = oracle.matmul(h, X)
    P = to_product_tree(gpu, h).repack()
    P.set_option(gpu.OPT_PIPELINE_COLS, 1 << 20)
    P.set_option(gpu.OPT_BUSH_KERNEL, 0)
    P.set_option(gpu.OPT_FLOW_KERNEL, 0)
    l0 = P.launch_count()
    Y0 = P @ X
    per_level = P.launch_count() - l0
    T0 = P.tmatmul(X)
    eligible = leafsize <= 64
    P.set_option(gpu.OPT_BUSH_KERNEL, 1 if eligible else 2)
    P.set_option(gpu.OPT_BUSH_LEVELS, levels)
    for _ in range(3):
        l0 = P.launch_count()
        Y1 = P @ X
        assert P.launch_count() - l0 <= 3 < per_level      # leaf-up, the bushes, leaf-down
        assert relerr(Y1, Y0) <= 1e-14 and (_ == 0 or np.array_equal(Y1, Yp))
        Yp = Y1
    assert P.get_option(gpu.OPT_BUSH_KERNEL) == 3
    assert relerr(Y1, ref) <= TOL
    l0 = P.launch_count()
    T1 = P.tmatmul(X)
    assert P.launch_count() - l0 <= 3
    assert relerr(T1, T0) <= 1e-14 and relerr(T1, oracle.matmul(oracle.adjoint(h), X)) <= TOL
    C0 = rng.standard_normal((n, nrhs))
    got = P.mul_(np.asfortranarray(C0.copy()), X, 0.7, -1.3)
    assert relerr(got, oracle.mul(C0.copy(), h, X, 0.7, -1.3)) <= TOL
    got = P.mul_(np.full((n, nrhs), np.nan, order="F"), X, 2.0, 0.0)   # beta == 0 never reads C
    assert np.isfinite(got).all() and relerr(got, 2.0 * ref) <= TOL
    X2 = rng.standard_normal((n, 2 * nrhs + 1))       # more column tiles than before: the flags grow
    assert relerr(P @ X2, oracle.matmul(h, X2)) <= TOL
    P.set_option(gpu.OPT_BUSH_LEVELS, 2 * 16 + 1)     # another cut of the same tree: plan rebuilt
    assert relerr(P @ X, Y0) <= 1e-14
    if not eligible:
        P.set_option(gpu.OPT_BUSH_KERNEL, 1)          # automatic: 128-row leaves stay on the dataflow kernel
        assert np.array_equal(P @ X, Y0) and P.get_option(gpu.OPT_BUSH_KERNEL) == 1
    P.close()


def test_graph_cache_with_fresh_pointers(gpu, oracle):
    """HSSB_OPT_USE_GRAPH with caller-owned device arrays: a repeating pointer set is captured once and replayed;
    fresh arrays on every call fall back to plain launches after two misses (no capture + instantiate per call), and a
    set that then repeats is captured again.  Same results either way."""
    import torch
    n, ls, r, k, seed = 8192, 128, 32, 64, 11
    h = oracle.synthetic_hss(n, ls, r, seed)
    X0 = oracle.synth_x(seed, n, k)
    ref = oracle.matmul(h, X0)
    st = torch.cuda.current_stream().cuda_stream
    with gpu.synthetic(n, ls, r, seed) as P:
        P.set_option(gpu.OPT_USE_GRAPH, 1)
        keep = []
        for i in range(8):                      # fresh X / Y every call
            X = torch.from_numpy(np.ascontiguousarray(X0.T)).cuda()
            Y = torch.full((k, n), float("nan"), dtype=torch.float64, device="cuda")
            keep.append((X, Y))
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
            torch.cuda.synchronize()
            assert relerr(Y.cpu().numpy().T, ref) <= TOL, i
        X, Y = keep[0]
        for i in range(4):                      # now a stable pair: captured on its second appearance, then replayed
            Y.fill_(float("nan"))
            P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
            torch.cuda.synchronize()
            assert relerr(Y.cpu().numpy().T, ref) <= TOL, i
