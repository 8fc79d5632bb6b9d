"""Device-resident time of BASELINE configs 1 (README Cauchy n=2001) and 2 (Cauchy n=2^16) — variable-rank trees,
any-shape kernel, CUDA-graph replay — next to the CPU restatement on the host cores."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import hssb200 as hb
import hss_oracle as o
from test_plan_cpu import to_product_tree

s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for name, n, leaf, tol, k in (("c1 README Cauchy n=2001", 2001, 64, 1e-6, 16), ("c2 Cauchy n=2^16", 2 ** 16, 64, 1e-9, 64)):
    if n <= 4096:
        h = o.hss(o.cauchy_matrix(n), leaf, tol, tol)
    else:
        cl = o.bisection_cluster(n, leaf)
        h = o.randcompress(o.cauchy_operator(n, device="cuda"), cl, cl, 30, tol, tol, rng=np.random.default_rng(16))
    t0 = time.perf_counter(); P = hb.pack(to_product_tree(hb, h)); tpack = time.perf_counter() - t0
    P.set_option(hb.OPT_USE_GRAPH, 1)
    Xh = np.random.default_rng(1).standard_normal((n, k))
    X = torch.from_numpy(np.ascontiguousarray(Xh.T)).cuda(); Y = torch.empty_like(X)
    for _ in range(3):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    C = np.empty((n, k)); o.mul(C, h, Xh)
    t0 = time.perf_counter()
    for _ in range(3):
        o.mul(C, h, Xh)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    err = np.linalg.norm(Y.cpu().numpy().T - C) / np.linalg.norm(C)
    fl, by = P.flops(k), P.algorithmic_bytes(k)
    print(f"{name}: rank {o.hssrank(h)}, pack {tpack * 1e3:.1f} ms, GPU {ms * 1e3:.1f} us ({fl / ms * 1e-6:.1f} GFLOP/s, {by / ms * 1e-6:.1f} GB/s), "
          f"CPU restatement {cpu_ms:.2f} ms ({fl / cpu_ms * 1e-6:.2f} GFLOP/s), rel.err {err:.1e}")
    P.close()

# per-phase breakdown of config 2 (profiled pass: plain launches + events between phases)
n, leaf, tol, k = 2 ** 16, 64, 1e-9, 64
cl = o.bisection_cluster(n, leaf)
h = o.randcompress(o.cauchy_operator(n, device="cuda"), cl, cl, 30, tol, tol, rng=np.random.default_rng(16))
P = hb.pack(to_product_tree(hb, h))
P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
acc = {}
for it in range(6):
    P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream); torch.cuda.synchronize()
    if it:
        for ph in P.phase_times():
            acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 5
print({k_: round(v * 1e3, 1) for k_, v in acc.items()}, "sum us", round(sum(acc.values()) * 1e3, 1))
