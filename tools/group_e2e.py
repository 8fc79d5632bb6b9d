"""One process, several GPUs (hssb_group_*): end-to-end time of `hssA * X` on whole host matrices against the
same matrix on one GPU.  python tools/group_e2e.py [log2 n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hssb200 as hb
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 21
n, ls, r, k, seed = 2 ** lg, 128, 32, 64, 3
nd = hb.device_count()
Xh = torch.randn((k, n), dtype=torch.float64).pin_memory()
Yh = torch.empty((k, n), dtype=torch.float64).pin_memory()
xs, ys = Xh.numpy().T, Yh.numpy().T
xp = np.array(xs, order="F", copy=True)
yp = np.empty((n, k), order="F")


def leg(P, x, y, steps=5):
    for _ in range(2):
        P.mul_(y, x)
    t = time.perf_counter()
    for _ in range(steps):
        P.mul_(y, x)
    return (time.perf_counter() - t) / steps * 1e3


ref = None
for P_ in [p for p in (1, 2, 4, 8) if p <= nd]:
    G = hb.synthetic_group(n, ls, r, seed, list(range(P_)))
    t_pin = leg(G, xs, ys)
    if ref is None:
        ref = np.array(ys[:, :2], copy=True)
    err = float(np.linalg.norm(ys[:, :2] - ref) / np.linalg.norm(ref))
    t_page = leg(G, xp, yp)
    flops = sum(sh.flops(k) for sh in G.shards)
    print(f"n = 2^{lg}, {P_} GPU(s) in one process: pinned {t_pin:.2f} ms ({flops / t_pin * 1e-9:.1f} TFLOP/s end to end), pageable {t_page:.2f} ms, "
          f"rel. diff to 1 GPU {err:.1e}", flush=True)
    G.close()
