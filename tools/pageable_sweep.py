"""End-to-end time of config 3 through hssb_matmul for pinned vs pageable caller memory.

Run once per HSSB_HOST_THREADS value (the worker pools are created once per process):
    for t in 2 4 8; do HSSB_HOST_THREADS=$t python tools/pageable_sweep.py; done
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hssb200 as hb

n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
P = hb.synthetic(n, ls, r, seed)


def leg(xs, ys, steps=5, fresh=False):
    for _ in range(2):
        P.mul_(ys, xs)
    t = time.perf_counter()
    for _ in range(steps):
        if fresh:
            ys = np.empty((n, k), order="F")  # like `similar(B, ...)` in matmul.jl:13: untouched pages every call
        P.mul_(ys, xs)
    return (time.perf_counter() - t) / steps * 1e3, ys


Xh = torch.randn((k, n), dtype=torch.float64).pin_memory()
Yh = torch.empty((k, n), dtype=torch.float64).pin_memory()
xs, ys = Xh.numpy().T, Yh.numpy().T
t_pin, _ = leg(xs, ys)
ref = np.array(ys[:, :2], copy=True)
xp = np.array(xs, order="F", copy=True)
yp = np.empty((n, k), order="F")
out = {"threads": os.environ.get("HSSB_HOST_THREADS", "default"), "cpus": os.cpu_count(), "pinned_ms": round(t_pin, 2)}
for name, mode in (("direct", 0), ("ring", 1)):
    P.set_option(hb.OPT_HOST_BOUNCE, mode)
    t, yy = leg(xp, yp)
    out[f"pageable_{name}_ms"] = round(t, 2)
    out[f"pageable_{name}_staged"] = P.get_option(hb.OPT_LAST_BOUNCE)
    out[f"pageable_{name}_err"] = float(np.linalg.norm(yy[:, :2] - ref) / np.linalg.norm(ref))
    t, yy = leg(xp, yp, fresh=True)
    out[f"pageable_{name}_fresh_y_ms"] = round(t, 2)
    yp[:] = 0
# mixed: pinned X, pageable Y and the other way round
P.set_option(hb.OPT_HOST_BOUNCE, 1)
out["pinX_pageY_ms"] = round(leg(xs, yp)[0], 2)
out["pageX_pinY_ms"] = round(leg(xp, ys)[0], 2)
# beta != 0 reads Y as well
yp[:] = 1.0
P.mul_(yp, xp, 2.0, 0.5)
out["alpha_beta_err"] = float(np.linalg.norm(yp[:, :2] - (2.0 * ref + 0.5)) / np.linalg.norm(ref))
print(out)
