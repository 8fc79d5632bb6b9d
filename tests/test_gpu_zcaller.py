"""The reference's in-library caller of the hot path through the C ABI:
randomized recompression of an operator that is itself an HSS matrix.  `randcompress_adaptive`
(src/compression.jl:311-356) samples `Scol = A*Ω`, `Srow = A'*Ω` (:326-327) and, per iteration, checks
`||Scol_test - hssA*Ω_test||` (:338-342) with nrhs = 20-30 — every one of those products runs on the GPU
here (forward, transposed, and on a freshly re-packed result); the compression arithmetic itself is the
oracle's restatement (CPU, not a GPU target)."""
import os

import numpy as np
import pytest

from test_plan_cpu import to_product_tree

pytestmark = pytest.mark.gpu


def test_randcompress_adaptive_caller(hb, oracle):
    if hb.device_count() == 0:
        pytest.skip("no B200 visible")
    n, leafsize, kest, bs, tol = 2048, 64, 20, 20, 1e-6
    A = oracle.cauchy_matrix(n)
    cl = oracle.bisection_cluster(n, leafsize)
    h0 = oracle.hss(A, leafsize, 1e-10, 1e-10)            # the operator: an accurate HSS form of A
    with hb.pack(to_product_tree(hb, h0)) as P0:

        class GpuOperator:                                  # LinearMap stand-in (src/linearmap.jl:31-32)
            shape = (n, n)

            def matmat(self, Om):                           # A*Ω   -> hssb_matmul
                return P0 @ Om

            def rmatmat(self, Om):                          # A'*Ω  -> hssb_matmul_t
                return P0.tmatmul(Om)

            def getindex(self, I, J):
                return A[np.ix_(I, J)]

        op = GpuOperator()
        Om = np.random.default_rng(5).standard_normal((n, kest + 10))
        assert np.linalg.norm(op.matmat(Om) - oracle.matmul(h0, Om)) <= 1e-12 * np.linalg.norm(oracle.matmul(h0, Om))
        ref_t = oracle.matmul(oracle.adjoint(h0), Om)
        assert np.linalg.norm(op.rmatmat(Om) - ref_t) <= 1e-12 * np.linalg.norm(ref_t)
        h1 = oracle.randcompress(op, cl, cl, kest, atol=tol, rtol=tol, rng=np.random.default_rng(3))   # :278-294
        # the error estimate of compression.jl:334-342, hssA*Ω_test on a freshly packed hssA
        Om_test = np.random.default_rng(4).standard_normal((n, bs))
        Scol_test = op.matmat(Om_test)
        with hb.pack(to_product_tree(hb, h1)) as P1:
            HOm = P1 @ Om_test
        nrm = np.sqrt(1.0 / bs) * np.linalg.norm(Scol_test)
        nrm_est = np.sqrt(1.0 / bs) * np.linalg.norm(Scol_test - HOm)
        failed = nrm_est > tol and nrm_est > tol * nrm        # :342
        assert not failed, (nrm_est, nrm)
        # same estimate with the CPU restatement of the product
        est_cpu = np.sqrt(1.0 / bs) * np.linalg.norm(oracle.matmul(h0, Om_test) - oracle.matmul(h1, Om_test))
        assert abs(nrm_est - est_cpu) <= 1e-3 * est_cpu + 1e-12 * nrm   # a difference of nearly equal vectors: rounding of the products shows
        assert np.linalg.norm(oracle.full(h1) - A) <= 50 * tol * np.linalg.norm(A)      # runtests.jl:39-40


def test_solve_golden_fixtures(hb, oracle):
    """tests/golden/solve/*.npz (dense solves, make_golden_solve.py) through the C ABI: hssb_solve."""
    if hb.device_count() == 0:
        pytest.skip("no B200 visible")
    import make_golden
    from test_ulv_cpu import solve_golden_files
    for f in solve_golden_files():
        z = np.load(f)
        h = make_golden.tree_from_npz(oracle, z)
        with hb.pack(to_product_tree(hb, h)) as P:
            got = P.solve(z["B"])
        assert np.linalg.norm(got - z["Z"]) <= 1e-13 * float(z["cond"]) * np.linalg.norm(z["Z"]), f


def test_ulv_fast_form_on_device(hb, oracle):
    """The ULV solve of a uniform tree on the product's fixed-shape kernels (HSSB_OPT_ULV_FAST, the default) against
    the general form on the any-shape kernel (option 0) and against the dense matrix."""
    if hb.device_count() == 0:
        pytest.fail("no B200 visible: the gpu-marked tests need the real device (no CPU fallback)")
    for n, ls, r, k in ((4096, 128, 32, 64), (4096, 128, 16, 20), (4096, 256, 32, 32)):
        with hb.synthetic(n, ls, r, 5) as P:
            B = oracle.synth_x(5, n, k)
            assert P.get_option(hb.OPT_ULV_FAST) == 2
            Z1 = P.solve(B)
            P.set_option(hb.OPT_ULV_FAST, 0)
            assert P.get_option(hb.OPT_ULV_FAST) == 0
            Z0 = P.solve(B)
            assert np.linalg.norm(Z1 - Z0) <= 1e-9 * np.linalg.norm(Z0)
            R = P @ Z1 - B
            A = oracle.full(oracle.synthetic_hss(n, ls, r, 5))
            assert np.linalg.norm(R) <= 1e-12 * np.linalg.norm(A, 2) * np.linalg.norm(Z1)
