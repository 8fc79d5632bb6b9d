"""Y = A' X on the config-3 matrix: forward plan over the adjoint twin pool (fixed-shape DMMA kernels)
against the any-shape transposed task table over the primary pool; also times building the twin."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
P = hb.synthetic(n, ls, r, seed)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 1)
X = torch.randn((k, n), dtype=torch.float64, device="cuda")
Y = {m: torch.empty_like(X) for m in ("fwd", "twin", "generic")}


def timed(name, trans, reps=10):
    for _ in range(3):
        P.matmul_dev(X.data_ptr(), n, Y[name].data_ptr(), n, k, stream=s.cuda_stream, trans=trans)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        P.matmul_dev(X.data_ptr(), n, Y[name].data_ptr(), n, k, stream=s.cuda_stream, trans=trans)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:8s} {ms:8.4f} ms  {P.flops(k) / ms * 1e-9:7.2f} TFLOP/s", flush=True)
    return ms


timed("fwd", False)
torch.cuda.synchronize(); t0 = time.perf_counter()
P.matmul_dev(X.data_ptr(), n, Y["twin"].data_ptr(), n, k, stream=s.cuda_stream, trans=True)
torch.cuda.synchronize()
print(f"first transposed product (builds the {P.info.pool_bytes * 1e-9:.2f} GB twin): {(time.perf_counter() - t0) * 1e3:.2f} ms; "
      f"option state {P.get_option(hb.OPT_ADJOINT_TWIN)}")
timed("twin", True)
P.set_option(hb.OPT_ADJOINT_TWIN, 0)
timed("generic", True)
err = (torch.linalg.norm(Y["twin"] - Y["generic"]) / torch.linalg.norm(Y["generic"])).item()
print(f"twin vs any-shape transposed plan: rel err {err:.3e}")
# adjoint identity <X2, A X1> = <A' X2, X1>
X2 = torch.randn_like(X)
P.set_option(hb.OPT_ADJOINT_TWIN, 1)
P.matmul_dev(X2.data_ptr(), n, Y["twin"].data_ptr(), n, k, stream=s.cuda_stream, trans=True)
torch.cuda.synchronize()
a, b = (X2 * Y["fwd"]).sum().item(), (Y["twin"] * X).sum().item()
print(f"adjoint identity: {a:.15e} vs {b:.15e}, rel {(abs(a - b) / (torch.linalg.norm(X2) * torch.linalg.norm(Y['fwd'])).item()):.3e}")
