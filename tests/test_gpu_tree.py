"""Persistent tree kernel (csrc/hssb_tree.cuh): every merge / translate level of src/matmul.jl:39 and
:52-56 in one cooperative launch with grid barriers.  Checked against the oracle AND, bit for bit,
against the one-launch-per-level schedule (same DMMA accumulation order per output element)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def gpu(hb):
    if hb.device_count() < 1:
        pytest.fail("no B200 visible: the gpu-marked tests need the real device (no CPU fallback)")
    return hb


SHAPES = [  # n, leafsize, rank
    (256, 128, 32),      # depth 1: a single root translate, no merges
    (512, 128, 16),
    (4096, 128, 32),
    (8192, 128, 64),
    (8192, 256, 64),
    (4096, 256, 16),
    (32768, 128, 32),    # 256 leaves: wide and narrow slice levels in one sweep
]


@pytest.mark.parametrize("n,ls,r", SHAPES)
def test_tree_kernel_matches_levels_and_oracle(gpu, oracle, n, ls, r):
    seed = 4242 + n + r
    h = oracle.synthetic_hss(n, ls, r, seed) if n <= 8192 else None
    with gpu.synthetic(n, ls, r, seed) as P:
        assert P.get_option(gpu.OPT_TREE_KERNEL) == 0   # measured slower than graph-replayed level launches: opt-in
        for k in (1, 9, 20, 33, 64, 65, 130, 200):
            X = oracle.synth_x(seed + k, n, k)
            P.set_option(gpu.OPT_TREE_KERNEL, 1)
            l0 = P.launch_count()
            Y1 = P @ X
            used = P.launch_count() - l0
            P.set_option(gpu.OPT_TREE_KERNEL, 0)
            l0 = P.launch_count()
            Y0 = P @ X
            per_level = P.launch_count() - l0
            assert np.array_equal(Y1, Y0), (n, ls, r, k)
            assert used <= per_level
            if h is not None:
                assert relerr(Y1, oracle.matmul(h, X)) <= TOL, (n, ls, r, k)
        # three launches per product: leaf-up, tree, leaf-down (host entry pipelines column blocks, so
        # count on a single-block call)
        P.set_option(gpu.OPT_TREE_KERNEL, 1)
        P.set_option(gpu.OPT_PIPELINE_COLS, 1 << 20)
        l0 = P.launch_count()
        P @ oracle.synth_x(seed, n, 64)
        assert P.launch_count() - l0 == 3


def test_tree_kernel_graph_replay_and_alpha_beta(gpu, oracle):
    import torch
    n, ls, r, k, seed = 16384, 128, 32, 64, 17
    with gpu.synthetic(n, ls, r, seed) as P:
        st = torch.cuda.current_stream().cuda_stream
        X = torch.zeros((k, n), dtype=torch.float64, device="cuda")
        gpu._check(gpu.lib().hssb_synthetic_rhs(seed, n, k, 0, n, X.data_ptr(), n, 0, st))
        Y0 = torch.randn((k, n), dtype=torch.float64, device="cuda")
        outs = []
        for tree, graph in ((0, 0), (1, 0), (1, 1), (2, 0)):
            P.set_option(gpu.OPT_TREE_KERNEL, tree)
            P.set_option(gpu.OPT_USE_GRAPH, graph)
            Y = Y0.clone()
            for _ in range(3):   # replays reuse the barrier words: the epoch bookkeeping must hold
                Y.copy_(Y0)
                P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, 0.75, -0.5, stream=st)
            torch.cuda.synchronize()
            outs.append(Y.cpu().numpy())
        for o in outs[1:]:
            assert np.array_equal(o, outs[0])
        h = oracle.synthetic_hss(n, ls, r, seed)
        Xh = X.cpu().numpy().T
        ref = oracle.mul(np.asfortranarray(Y0.cpu().numpy().T.copy()), h, Xh, 0.75, -0.5)
        assert relerr(outs[0].T, ref) <= TOL


def test_tree_kernel_adjoint_twin(gpu, oracle):
    n, ls, r, k, seed = 4096, 128, 32, 40, 5
    h = oracle.synthetic_hss(n, ls, r, seed)
    X = oracle.synth_x(seed, n, k)
    with gpu.synthetic(n, ls, r, seed) as P:
        assert relerr(P.tmatmul(X), oracle.matmul(oracle.adjoint(h), X)) <= TOL
        assert P.get_option(gpu.OPT_ADJOINT_TWIN) == 2
