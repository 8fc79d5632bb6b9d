import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import hssb200 as hb
import hss_oracle as o
n, ls, r, seed = 8192, 128, 32, 23
ks = [int(a) for a in sys.argv[1:]] or [1, 20, 64, 130]
h = o.synthetic_hss(n, ls, r, seed)
G = hb.synthetic_group(n, ls, r, seed, [0, 0])
single = hb.synthetic(n, ls, r, seed) if os.environ.get("WITH_SINGLE") else None
for k in ks:
    X = o.synth_x(seed + k, n, k)
    t = time.time()
    try:
        Y = G @ X
        if single is not None:
            Ys = single @ X
        print("k", k, "ok", np.linalg.norm(Y - o.matmul(h, X)) / np.linalg.norm(Y), round(time.time() - t, 2), flush=True)
    except Exception as e:
        print("k", k, "FAILED after", round(time.time() - t, 2), str(e)[:200], flush=True)
        break
