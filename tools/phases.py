"""Per-phase device time (profiled pass: plain launches + events) of a synthetic matrix: python tools/phases.py n leafsize rank nrhs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hssb200 as hb
n, ls, r, k = [int(a) for a in sys.argv[1:5]]
opts = [a.split("=") for a in sys.argv[5:]]
P = hb.synthetic(n, ls, r, 3)
for o, v in opts:
    P.set_option(getattr(hb, o), int(v))
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
P.set_option(hb.OPT_USE_GRAPH, 0); P.set_option(hb.OPT_PROFILE, 1)
X = torch.randn((k, n), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
acc = {}
tasks = {}
for it in range(6):
    P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=s.cuda_stream); torch.cuda.synchronize()
    if it:
        for ph in P.phase_times():
            acc[ph["name"]] = acc.get(ph["name"], 0.0) + ph["ms"] / 5
            tasks[ph["name"]] = ph["ntasks"]
L = n // ls
for a, v in acc.items():
    if a.startswith("leaf"):
        continue
    # per task: two r x r generator blocks, two operand tiles in, one out (padded leading dimension r + 4)
    byts = tasks[a] * 8 * (r + 4) * (2 * r + 3 * k)
    fl = tasks[a] * 4 * r * r * k
    print("%-14s tasks %6d  %8.1f us   %6.2f TB/s  %6.2f TF/s" % (a, tasks[a], v * 1e3, byts / v * 1e-9, fl / v * 1e-9))
print("leaf_up %.1f us leaf_down %.1f us levels %.1f us total %.1f us" % (acc["leaf_up"] * 1e3, acc["leaf_down"] * 1e3, (sum(acc.values()) - acc["leaf_up"] - acc["leaf_down"]) * 1e3, sum(acc.values()) * 1e3))
