"""ULV solver (hssA \\ B) on the config-3 matrix: device factorisation time, solve time per call, the
backward error of A (A \\ B) = B with ||A||_2 from power iterations on the device, and the CPU
restatement of ulvfactor.jl (which, like the reference, factorises on every call) on a smaller n."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import hssb200 as hb
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 2 ** 20
FAST = "--fast" in sys.argv   # experimental HSSB_OPT_ULV_FAST: fixed-shape kernels on the solve plan
ls, r, k, seed = 128, 32, 64, 3
P = hb.synthetic(n, ls, r, seed)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 1)
if FAST:
    P.set_option(hb.OPT_ULV_FAST, 1)
    print("fast form:", P.get_option(hb.OPT_ULV_FAST) == 2)
ui = P.ulv_info
print(f"n {n}: factor pool {ui.pool_bytes * 1e-9:.2f} GB (generators {P.info.pool_bytes * 1e-9:.2f} GB), solve flops/rhs {ui.flops_per_rhs:.3e} (product {P.info.flops_per_rhs:.3e})")
torch.cuda.synchronize(); t0 = time.perf_counter()
P.ulv_factor()
torch.cuda.synchronize(); tf = time.perf_counter() - t0
print(f"factorisation: {tf * 1e3:.1f} ms (device time of its kernels: {P.get_option(hb.OPT_LAST_FACTOR_US) * 1e-3:.1f} ms)")
B = torch.randn((k, n), dtype=torch.float64, device="cuda"); Z = torch.empty_like(B); Y = torch.empty_like(B)
for _ in range(3):
    P.solve_dev(B.data_ptr(), n, Z.data_ptr(), n, k, stream=s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    P.solve_dev(B.data_ptr(), n, Z.data_ptr(), n, k, stream=s.cuda_stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = ui.flops_per_rhs * k
by = ui.flops_per_rhs / 2 * 8 + 2 * 8 * n * k
print(f"solve nrhs {k}: {ms:.3f} ms  {fl / ms * 1e-9:.2f} TFLOP/s  {by / ms * 1e-6:.0f} GB/s (factor matrices read once + B + Z)")
# per-phase device times of the solve (plain launches + events)
P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
acc = {}
for it in range(4):
    P.solve_dev(B.data_ptr(), n, Z.data_ptr(), n, k, stream=s.cuda_stream); torch.cuda.synchronize()
    if it:
        for ph in P.phase_times():
            a = acc.setdefault(ph["name"], [0.0, ph["flops_per_rhs"] * k, ph["ntasks"]])
            a[0] += ph["ms"] / 3
big = {nm: v for nm, v in acc.items() if v[0] > 0.03 * ms}
print("solve phases (ms, TFLOP/s, tasks):", {nm: (round(v[0], 3), round(v[1] / v[0] * 1e-9, 1), v[2]) for nm, v in big.items()},
      "rest", round(sum(v[0] for nm, v in acc.items() if nm not in big), 3), "ms in", len(acc) - len(big), "phases")
P.set_option(hb.OPT_PROFILE, 0); P.set_option(hb.OPT_USE_GRAPH, 1)
# backward error: ||B - A Z|| / (||A||_2 ||Z||), ||A||_2 by power iteration with A and A'
P.matmul_dev(Z.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
torch.cuda.synchronize()
res = torch.linalg.norm(Y - B).item()
v = torch.randn((1, n), dtype=torch.float64, device="cuda"); w = torch.empty_like(v)
for _ in range(30):
    v /= torch.linalg.norm(v)
    P.matmul_dev(v.data_ptr(), n, w.data_ptr(), n, 1, stream=s.cuda_stream)
    P.matmul_dev(w.data_ptr(), n, v.data_ptr(), n, 1, stream=s.cuda_stream, trans=True)
    torch.cuda.synchronize()
norm2 = torch.linalg.norm(w).item()
print(f"residual ||B - A Z|| / ||B|| = {res / torch.linalg.norm(B).item():.3e}; backward error / (||A||_2 ||Z||) = {res / (norm2 * torch.linalg.norm(Z).item()):.3e}  (||A||_2 ~ {norm2:.3e})")
# CPU restatement on a smaller matrix (factorises + solves in one pass, like the reference)
import hss_oracle as o, hss_ulv_oracle as uo
nc = min(n, 2 ** 15)
h = o.synthetic_hss(nc, ls, r, seed)
Bc = np.random.default_rng(0).standard_normal((nc, k))
t0 = time.perf_counter(); uo.ulvfactsolve(h, Bc); tc = time.perf_counter() - t0
print(f"CPU restatement of ulvfactsolve at n = {nc}: {tc:.2f} s  -> x{n // nc} leaves = {tc * n / nc:.1f} s at n = {n}")
