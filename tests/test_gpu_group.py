"""One process, one call, several devices (hssb_group_*, SURVEY 8b: "a single ccall from one Julia thread
drives all GPUs").  Shards may share a device, so the sharded plan, the replicated top tree and the
peer-store exchange kernels are exercised on a single-GPU box as well; with two or more GPUs visible the
same tests also run across devices (NVLink peer stores, no IPC, no NCCL)."""
import numpy as np
import pytest

from test_plan_cpu import to_product_tree

pytestmark = pytest.mark.gpu
TOL = 1e-12


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def gpu(hb):
    if hb.device_count() < 1:
        pytest.fail("no B200 visible: the gpu-marked tests need the real device (no CPU fallback)")
    return hb


def device_sets(hb, P):
    """Device lists for P shards on this box (at most two shards may share a GPU)."""
    nd = hb.device_count()
    sets = []
    if P <= 2:
        sets.append([0] * P)                           # both shards on one GPU: runs on every box
    if nd >= 2 and 2 * nd >= P:
        sets.append([g % nd for g in range(P)])        # spread over the GPUs of the box
    if not sets:
        pytest.skip(f"{P} shards need at least {(P + 1) // 2} GPUs, {nd} visible")
    return sets


@pytest.mark.parametrize("P", [2, 4])
def test_group_ragged_tree_host_entry(gpu, oracle, P):
    rng = np.random.default_rng(100 + P)
    n, k = 1111, 7
    cl = oracle.bisection_cluster(n, 60)
    h = oracle.random_hss(cl, cl, rng, 1, 9)
    X = rng.standard_normal((n, k))
    ref = oracle.matmul(h, X)
    C0 = rng.standard_normal((n, k))
    ref_ab = oracle.mul(C0.copy(), h, X, 0.7, -1.3)
    tree = to_product_tree(gpu, h)
    for devs in device_sets(gpu, P):
        with gpu.pack_group(tree, devs) as G:
            assert G.size == P and (G.m, G.n) == (n, n)
            assert sum(sh.info.local_m for sh in G.shards) == n
            Y = G @ X                                   # `*(hssA, B)`, matmul.jl:13, on P shards
            assert relerr(Y, ref) <= TOL, devs
            assert relerr(G @ X[:, 0], ref[:, 0]) <= TOL
            got = G.mul_(np.asfortranarray(C0.copy()), X, 0.7, -1.3)      # mul!, matmul.jl:18
            assert relerr(got, ref_ab) <= TOL
            got = G.mul_(np.full((n, k), np.nan, order="F"), X, 2.0, 0.0)  # beta == 0 never reads C
            assert np.isfinite(got).all() and relerr(got, 2.0 * ref) <= TOL
            with pytest.raises(gpu.DimensionMismatch):
                G.mul_(np.zeros((n, k), order="F"), X[:n - 1])
            with pytest.raises(gpu.DimensionMismatch):
                G.mul_(np.zeros((n - 1, k), order="F"), X)
            for _ in range(3):                          # repeated calls: exchange epochs / acknowledgements hold
                assert relerr(G @ X, ref) <= TOL


@pytest.mark.parametrize("P", [2, 4, 8])
def test_group_synthetic_fixed_shape_kernels(gpu, oracle, P):
    n, ls, r, seed = 8192, 128, 32, 23
    h = oracle.synthetic_hss(n, ls, r, seed)
    for devs in device_sets(gpu, P):
        with gpu.synthetic_group(n, ls, r, seed, devs) as G, gpu.synthetic(n, ls, r, seed) as single:
            for k in (1, 20, 64, 130):
                X = oracle.synth_x(seed + k, n, k)
                ref = oracle.matmul(h, X)
                Y = G @ X
                assert relerr(Y, ref) <= TOL, (devs, k)
                assert relerr(Y, single @ X) <= 1e-14
            X = oracle.synth_x(seed, n, 40)
            assert relerr(G.tmatmul(X), oracle.matmul(oracle.adjoint(h), X)) <= TOL   # A' X: adjoint twin pools
            A = np.random.default_rng(1).standard_normal((3, n))
            assert relerr(A @ G, A @ oracle.full(h)) <= TOL                            # `*(A, hssB)`, matmul.jl:14


def test_group_device_pointers_and_graph_replay(gpu, oracle):
    import torch
    n, ls, r, seed, k, P = 16384, 128, 32, 29, 64, 2
    h = oracle.synthetic_hss(n, ls, r, seed)
    X = oracle.synth_x(seed, n, k)
    ref = oracle.matmul(h, X)
    for devs in device_sets(gpu, P):
        with gpu.synthetic_group(n, ls, r, seed, devs) as G:
            G.set_option(gpu.OPT_USE_GRAPH, 1)
            dX, dY = [], []
            for g, sh in enumerate(G.shards):
                dev = torch.device("cuda", devs[g])
                r0, rows = sh.info.local_col0, sh.info.local_n
                dX.append(torch.from_numpy(np.ascontiguousarray(X[r0:r0 + rows].T)).to(dev))   # (k, rows) = column-major rows x k
                dY.append(torch.full((k, sh.info.local_m), float("nan"), dtype=torch.float64, device=dev))
            for d in set(devs):
                torch.cuda.synchronize(d)
            rows = n // P
            for _ in range(3):
                G.matmul_dev([t.data_ptr() for t in dX], rows, [t.data_ptr() for t in dY], rows, k)
            G.sync()
            Y = np.vstack([t.cpu().numpy().T for t in dY])
            assert relerr(Y, ref) <= TOL, devs


def test_group_of_one_is_a_plain_handle(gpu, oracle):
    n, ls, r, seed = 4096, 128, 16, 31
    h = oracle.synthetic_hss(n, ls, r, seed)
    X = oracle.synth_x(seed, n, 9)
    with gpu.synthetic_group(n, ls, r, seed, [0]) as G:
        assert G.size == 1 and relerr(G @ X, oracle.matmul(h, X)) <= TOL
    with pytest.raises(gpu.HssbError):
        gpu.synthetic_group(n, ls, r, seed, [0, 0, 0])        # not a power of two
    with pytest.raises(gpu.HssbError):
        gpu.synthetic_group(n, ls, r, seed, [0, 0, 0, 0])     # more than two shards on one device
    with pytest.raises(gpu.HssbError):
        gpu.synthetic_group(128, 128, r, seed, [0, 0])        # tree too shallow for two shards


def test_group_pipelined_host_entry_pageable_and_pinned(gpu, oracle):
    """A call large enough to be pipelined over column blocks (every block carries its own exchange) and to be
    staged through the pinned rings (pageable numpy memory), two shards in one process: against one GPU, bit for bit."""
    import torch
    n, ls, r, seed, k = 2 ** 17, 128, 32, 37, 64
    X = oracle.synth_x(seed, n, k)
    with gpu.synthetic(n, ls, r, seed) as single:
        ref = single @ X
    lazy = oracle.LazySyntheticHss(n, ls, r, seed)
    for devs in device_sets(gpu, 2):
        with gpu.synthetic_group(n, ls, r, seed, devs) as G:
            Y = np.full((n, k), np.nan, order="F")
            G.mul_(Y, X)                                    # pageable: rings + worker threads, 8 column blocks
            assert all(sh.get_option(gpu.OPT_LAST_BOUNCE) == 3 for sh in G.shards)
            assert np.array_equal(Y, ref), devs
            Xh = torch.empty((k, n), dtype=torch.float64).pin_memory()
            Yh = torch.empty((k, n), dtype=torch.float64).pin_memory()
            Xh.numpy().T[:] = X
            G.mul_(Yh.numpy().T, Xh.numpy().T)              # pinned: straight to the copy engines
            assert all(sh.get_option(gpu.OPT_LAST_BOUNCE) == 0 for sh in G.shards)
            assert np.array_equal(Yh.numpy().T, ref)
    # and the single-GPU result itself against the oracle on a few leaves (sparse-support right-hand side)
    s_lo, s_len = 3 * ls + 5, 2 * ls
    Xs = np.random.default_rng(0).standard_normal((s_len, 4))
    Xsp = np.zeros((n, 4))
    Xsp[s_lo:s_lo + s_len] = Xs
    with gpu.synthetic_group(n, ls, r, seed, device_sets(gpu, 2)[-1]) as G:
        Ysp = G @ Xsp
    for lo, yref in lazy.rows([0, 3 * ls, (n // ls // 2) * ls, n - ls], s_lo, Xs).items():
        assert relerr(Ysp[lo:lo + ls], yref) <= TOL, lo
