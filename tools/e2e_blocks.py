"""End-to-end (host pinned X -> Y) time of config 3 vs the column-block size of the pipelined host entry."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
P = hb.synthetic(n, ls, r, seed)
Xh = torch.randn((k, n), dtype=torch.float64).pin_memory()
Yh = torch.empty((k, n), dtype=torch.float64).pin_memory()
xs, ys = Xh.numpy().T, Yh.numpy().T
for cols in (64, 32, 16, 8, 4):
    P.set_option(hb.OPT_PIPELINE_COLS, cols)
    for _ in range(2):
        P.mul_(ys, xs)
    t = time.perf_counter()
    for _ in range(5):
        P.mul_(ys, xs)
    print("block of", cols, "columns:", round((time.perf_counter() - t) / 5 * 1e3, 3), "ms per product")
