"""Device-resident time of one product on the config-3 matrix for several nrhs (column-tile widths 32/64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, seed = 2 ** 20, 128, 32, 3
P = hb.synthetic(n, ls, r, seed)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 1)
dmma = hb.measure_peak(1, 20000)
for k in (1, 8, 20, 32, 48, 64, 96, 128, 256):
    X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    for _ in range(3):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl, by = P.flops(k), P.algorithmic_bytes(k)
    t_roof = max(fl / (dmma * 1e12), by / 6550.7e9) * 1e3
    print(f"nrhs {k:4d}: {ms:8.4f} ms  {fl / ms * 1e-9:8.2f} TFLOP/s  {by / ms * 1e-6:8.1f} GB/s  roofline {t_roof:.4f} ms -> {t_roof / ms:.3f}")
