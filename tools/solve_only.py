"""One ULV solve on the config-3 matrix with plain launches (for ncu: -k regex:generic_level_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, k, seed = 2 ** 20, 128, 32, 64, 3
P = hb.synthetic(n, ls, r, seed)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 0)
P.ulv_factor()
B = torch.randn((k, n), dtype=torch.float64, device="cuda"); Z = torch.empty_like(B)
for _ in range(2):
    P.solve_dev(B.data_ptr(), n, Z.data_ptr(), n, k, stream=s.cuda_stream)
torch.cuda.synchronize()
print("done", P.launch_count())
