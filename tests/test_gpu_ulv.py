"""GPU parity tests of the ULV solver (hssA \\ B, src/ulvfactor.jl:10-107) through
the C ABI: device factorisation (hssb_ulv_factor) + level-scheduled solve (hssb_solve) against the oracle
restatement of ulvfactor.jl, the dense solve and the host instantiation of the same node routine."""
import numpy as np
import pytest

from test_plan_cpu import to_product_tree
from test_ulv_cpu import CASES, shifted

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(hb):
    if hb.device_count() == 0:
        pytest.skip("no B200 visible")
    return hb


@pytest.mark.parametrize("n,leafsize,nrhs,rmin,rmax", CASES + [(4000, 128, 70, 5, 40), (1500, 256, 33, 17, 17)])
def test_solve_matches_oracle(gpu, oracle, ulv_oracle, n, leafsize, nrhs, rmin, rmax):
    rng = np.random.default_rng(n + 13 * leafsize)
    cl = oracle.bisection_cluster(n, leafsize)
    h = shifted(oracle, oracle.random_hss(cl, cl, rng, rmin, rmax), 4.0 * np.sqrt(leafsize))
    A = oracle.full(h)
    B = rng.standard_normal((n, nrhs))
    ref = ulv_oracle.ulvfactsolve(h, B)
    cond = np.linalg.cond(A)
    tree = to_product_tree(gpu, h)
    with gpu.pack(tree) as P:
        assert P.ulv_info.supported == 1 and P.ulv_info.factored == 0
        Z = P.solve(B)                       # factorises on first use
        assert P.ulv_info.factored == 1
        # backward error (scale free) and forward parity with the oracle (conditioning dependent)
        assert np.linalg.norm(A @ Z - B) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Z)
        assert np.linalg.norm(Z - ref) <= 1e-14 * cond * np.linalg.norm(ref) + 1e-300
        # vector right-hand side, repeated solve (graph replay), product still right afterwards
        z1 = P.solve(B[:, 0])
        assert z1.shape == (n,) and np.linalg.norm(z1 - Z[:, 0]) <= 1e-13 * np.linalg.norm(z1) + 1e-300
        X = rng.standard_normal((n, 2))
        assert np.linalg.norm(P @ X - A @ X) <= 1e-12 * np.linalg.norm(A @ X)
        # the device factor pool equals the host instantiation of the same node routine (up to fma contraction)
        dev = P.debug_ulv_pool()
    Zt = gpu.ulvfactsolve(tree, B)           # the reference-named entry on the tree (packs on first use)
    assert np.linalg.norm(Zt - Z) <= 1e-13 * np.linalg.norm(Z) + 1e-300
    tree._packed.close()
    with gpu.pack(tree, plan_only=True) as Q:
        host = Q.debug_ulv_pool(factor_on_host=True)
    assert np.linalg.norm(dev - host) <= 1e-11 * np.linalg.norm(host)


def test_reference_solver_assertion_restated(gpu, oracle, ulv_oracle):
    """test/runtests.jl:66-67 through the C ABI: x = hssA \\ rhs against x0 = A \\ rhs on the reference's test
    matrix (compress at tol 1e-6, leafsize 50): ||x0 - x|| / ||x0|| <= 50 * tol, and parity with the oracle."""
    from test_ulv_cpu import reference_test_matrix
    A = reference_test_matrix()
    tol, c = 1e-6, 50.0
    h = oracle.hss(A, leafsize=50, atol=tol, rtol=tol)
    rhs = np.random.default_rng(1).standard_normal((2001, 5))
    x0 = np.linalg.solve(A, rhs)
    xr = ulv_oracle.ulvfactsolve(h, rhs)
    with gpu.pack(to_product_tree(gpu, h)) as P:
        x = P.solve(rhs)
    assert np.linalg.norm(x0 - x) / np.linalg.norm(x0) <= c * tol
    assert np.linalg.norm(xr - x) / np.linalg.norm(xr) <= 1e-12


def test_solve_then_multiply_uniform(gpu, oracle, ulv_oracle):
    """Uniform synthetic trees (padded pool): A (A \\ B) == B with the product path, factor ahead of time."""
    import torch
    for n, ls, r, k, seed in ((4096, 128, 32, 64, 5), (4096, 256, 64, 20, 6)):
        with gpu.synthetic(n, ls, r, seed) as P:
            P.ulv_factor()
            assert P.ulv_info.factored == 1
            B = oracle.synth_x(seed, n, k)
            Z = P.solve(B)
            R = P @ Z - B
            # scale-free backward error with ||A||_2 from the dense expansion at this size
            A = oracle.full(oracle.synthetic_hss(n, ls, r, seed))
            assert np.linalg.norm(R) <= 1e-12 * np.linalg.norm(A, 2) * np.linalg.norm(Z)
            if ls == 128:
                ref = ulv_oracle.ulvfactsolve(oracle.synthetic_hss(n, ls, r, seed), B)
                assert np.linalg.norm(Z - ref) <= 1e-14 * np.linalg.cond(A) * np.linalg.norm(ref)
            # device entry on torch memory with padded leading dimensions
            ldb, ldz = n + 8, n + 24
            Bd = torch.zeros((k, ldb), dtype=torch.float64, device="cuda")
            Bd[:, :n] = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
            Zd = torch.full((k, ldz), float("nan"), dtype=torch.float64, device="cuda")
            torch.cuda.synchronize()
            P.solve_dev(Bd.data_ptr(), ldb, Zd.data_ptr(), ldz, k, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            Zh = Zd.cpu().numpy()
            assert np.linalg.norm(Zh[:, :n].T - Z) <= 1e-13 * np.linalg.norm(Z)
            assert np.isnan(Zh[:, n:]).all()


def test_full_size_config3_solve(gpu, oracle):
    """The config-3 matrix (n = 2^20, leaf 128, rank 32) at full size: factorise on the device, solve 64
    right-hand sides, multiply back with the product path.  Backward error ||B - A Z|| / (||A||_2 ||Z||)
    with ||A||_2 from power iterations (A and A' products on the device)."""
    import torch
    n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
    with gpu.synthetic(n, ls, r, seed) as P:
        st = torch.cuda.current_stream().cuda_stream
        B = torch.randn((k, n), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
        Z, Y = torch.empty_like(B), torch.empty_like(B)
        P.solve_dev(B.data_ptr(), n, Z.data_ptr(), n, k, stream=st)
        P.matmul_dev(Z.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
        torch.cuda.synchronize()
        assert P.ulv_info.factored == 1 and bool(torch.isfinite(Z).all())
        v = torch.randn((1, n), dtype=torch.float64, device="cuda")
        w = torch.empty_like(v)
        for _ in range(20):
            v /= torch.linalg.norm(v)
            P.matmul_dev(v.data_ptr(), n, w.data_ptr(), n, 1, stream=st)
            P.matmul_dev(w.data_ptr(), n, v.data_ptr(), n, 1, stream=st, trans=True)
            torch.cuda.synchronize()
        norm2 = torch.linalg.norm(w).item()     # a lower bound of ||A||_2, hence a pessimistic backward error
        eta = torch.linalg.norm(Y - B).item() / (norm2 * torch.linalg.norm(Z).item())
        assert eta <= 1e-12, eta


def test_solver_errors(gpu, oracle):
    rng = np.random.default_rng(3)
    rcl = oracle.bisection_cluster(500, 70)
    ccl = oracle.bisection_cluster(333, 47)
    with gpu.pack(to_product_tree(gpu, oracle.random_hss(rcl, ccl, rng, 1, 7))) as P:   # not square
        assert P.ulv_info.supported == 0
        with pytest.raises(gpu.HssbError):
            P.solve(np.zeros((500, 1)))
    with gpu.synthetic(2048, 128, 32, 1) as S:
        with pytest.raises(gpu.DimensionMismatch):
            S.solve(np.zeros((2047, 2)))


def test_singular_matrix_on_device(gpu, oracle):
    """A zero pivot in the device factorisation surfaces as SingularException (HSSB_ERR_SINGULAR), like `D \\ b` at
    ulvfactor.jl:83; no factors are cached, and the product of the same handle keeps working."""
    rng = np.random.default_rng(4)
    cl = oracle.bisection_cluster(256, 32)
    h = oracle.random_hss(cl, cl, rng, 2, 5)

    def zero_d(t):
        if t.leafnode:
            t.D[...] = 0.0
            t.U[...] = 0.0
        else:
            zero_d(t.A11)
            zero_d(t.A22)

    zero_d(h)   # A = 0: every reduced block is singular
    X = rng.standard_normal((256, 3))
    with gpu.pack(to_product_tree(gpu, h)) as P:
        with pytest.raises(gpu.SingularException):
            P.solve(X)
        assert P.ulv_info.factored == 0
        with pytest.raises(gpu.SingularException):
            P.ulv_factor()
        assert np.linalg.norm(P @ X) == 0.0


@pytest.mark.parametrize("n,ls,r,k", [(4096, 128, 32, 9), (2048, 128, 16, 3), (4096, 256, 32, 40)])
def test_right_division_by_the_adjoint_solve(gpu, oracle, ulv_oracle, n, ls, r, k):
    """`/(A, hssB) = ulvfactsolve(hssB', collect(A'))'` (hssmatrix.jl:236) without building hssB': on a uniform tree
    A' has the shapes of A, so hssb_solve_t shares the solve plan and factorises the adjoint twin pool."""
    seed = 40 + r
    h = oracle.synthetic_hss(n, ls, r, seed)
    A = oracle.full(h)
    B = oracle.synth_x(seed, n, k)
    with gpu.synthetic(n, ls, r, seed) as P:
        Zt = P.solve_t(B)                                   # A' \ B
        assert np.linalg.norm(A.T @ Zt - B) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Zt)
        ref = ulv_oracle.ulvfactsolve(oracle.adjoint(h), B)  # the reference's route: factorise the adjoint copy
        assert np.linalg.norm(Zt - ref) <= 1e-14 * np.linalg.cond(A) * np.linalg.norm(ref)
        M = B.T[:5].copy()                                  # a 5 x n matrix: M / hssB
        Q = M / P
        assert Q.shape == M.shape and np.linalg.norm(Q @ A - M) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Q)
        Z = P.solve(B)                                      # the forward solve keeps its own factors
        assert np.linalg.norm(A @ Z - B) <= 1e-13 * np.linalg.norm(A, 2) * np.linalg.norm(Z)
        assert np.linalg.norm(P.solve_t(B) - Zt) == 0.0     # second call: cached factors, graph replay
    rng = np.random.default_rng(1)
    cl = oracle.bisection_cluster(300, 40)
    with gpu.pack(to_product_tree(gpu, shifted(oracle, oracle.random_hss(cl, cl, rng, 1, 5), 20.0))) as G:
        with pytest.raises(gpu.HssbError):                  # ragged tree: no twin pool
            G.solve_t(np.zeros((300, 1)))
