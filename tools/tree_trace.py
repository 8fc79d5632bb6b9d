"""Per-level device time of the persistent tree kernel (csrc/hssb_tree.cuh) next to the per-level launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hssb200 as hb

cfgs = {"c3": (2 ** 20, 128, 32, 64), "c4": (2 ** 22, 128, 64, 128), "c5": (2 ** 24, 256, 64, 32), "c3k20": (2 ** 20, 128, 32, 20)}
for name in (sys.argv[1:] or ["c3"]):
    n, ls, r, k = cfgs[name]
    with hb.synthetic(n, ls, r, 3) as P:
        P.reserve(k)
        us = P.tree_trace(k)
        names = [p["name"] for p in P.phase_times() if p["kind"] not in (0, 4)]
        tasks = [p["ntasks"] for p in P.phase_times() if p["kind"] not in (0, 4)]
        print(name, "tree total us", round(sum(us), 2))
        for nm, t, u in zip(names, tasks, us):
            print(f"  {nm:16s} tasks {t:6d}  {u:8.2f} us")
