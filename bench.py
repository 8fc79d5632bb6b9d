#!/usr/bin/env python
"""bench.py — HSS x dense product throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config c3|c4|c5] [--impl reference]

A step = one `hssA * X` (reference src/matmul.jl:13-62) over one synthetic
right-hand side.  N = 1 runs BASELINE config 3 (n = 2^20, leafsize 128, rank 32,
nrhs 64: the configuration the north-star target is quoted on; configs[1], the
compressed Cauchy matrix, is a parity-test case).  N > 1 (launched by torchrun,
one rank per GPU) shards the tree by subtree: weak scaling, every GPU owns a
config-3-sized subtree (n = N * 2^20), one NCCL all-gather of the subtree-root Z
blocks per product.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

CONFIGS = {  # name: (n per GPU or total, leafsize, rank, nrhs, scaling)
    "c3": dict(n=2 ** 20, leafsize=128, rank=32, nrhs=64, scaling="weak",
               desc="synthetic random-generator HSS n=2^20 per GPU, leafsize 128, rank 32, nrhs 64"),
    "c4": dict(n=2 ** 22, leafsize=128, rank=64, nrhs=128, scaling="strong",
               desc="synthetic random-generator HSS n=2^22, leafsize 128, rank 64, nrhs 128, subtree-sharded"),
    "c5": dict(n=2 ** 24, leafsize=256, rank=64, nrhs=32, scaling="strong",
               desc="synthetic random-generator HSS n=2^24, leafsize 256, rank 64, nrhs 32, subtree-sharded"),
}
SEED = 3


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md).  Polls NVML
    (nvidia_ml_py) every ~2 ms from a thread, because the timed region is only tens of
    milliseconds long and `nvidia-smi -lms` cannot sample that fast; falls back to one
    nvidia-smi query if NVML is unavailable."""

    def __init__(self, gpu_index):
        self.samples, self.stop_flag, self.thread, self.nv, self.h = [], False, None, None, None
        self.idx = gpu_index
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), clk, reasons, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if not self.nv:
            out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
            if shutil.which("nvidia-smi"):
                try:
                    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                        "--format=csv,noheader,nounits", "-i", str(self.idx)], capture_output=True, text=True, timeout=10)
                    f = [x.strip() for x in q.stdout.strip().split(",")]
                    out.update(sm_mhz=float(f[0]), sm_max_mhz=float(f[1]), note="single nvidia-smi sample after the timed region")
                except Exception:
                    pass
            return out
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed region"
        if len(inside) < 3:  # very short region: widen to everything sampled under load (warm-up included)
            inside, window = self.samples, "warm-up + timed region"
        clks = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        return {"sm_mhz": clks[len(clks) // 2] if clks else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, v in names.items() if bits & v), "power_w_max": max((s[3] for s in inside), default=None),
                "samples": len(inside), "window": window, "source": "NVML polled every 2 ms"}


# ---------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference recursion (numpy -> OpenBLAS,
# one dgemm per node like src/matmul.jl), on a BOUNDED sample of the workload.
# ---------------------------------------------------------------------------
def cpu_sample(cfg, steps, warmup, sample_rows):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import hss_oracle as o
    ls, r, k = cfg["leafsize"], cfg["rank"], cfg["nrhs"]
    # a depth-d subtree of the full matrix, treated as root (rooted(), matmul.jl:24):
    # same leaves, same per-node GEMM shapes, 1/2^d of the work.
    h = o.synthetic_hss(sample_rows, ls, r, SEED)
    X = o.synth_x(SEED, sample_rows, k)
    C = np.empty((sample_rows, k))
    _, flops = o.algorithmic_counts(h, k)
    for _ in range(max(warmup, 1)):
        o.mul(C, h, X, 1.0, 0.0)
    ts = []
    for _ in range(max(steps, 1)):
        t = time.perf_counter()
        o.mul(C, h, X, 1.0, 0.0)  # copying slices, as matmul.jl:37-38
        ts.append(time.perf_counter() - t)
    best, mean = min(ts), sum(ts) / len(ts)
    try:
        from threadpoolctl import threadpool_info
        nthreads = max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        nthreads = os.cpu_count() or 1
    return dict(gflops=flops / mean * 1e-9, best_gflops=flops / best * 1e-9, ms=mean * 1e3, cores=nthreads,
                flops=flops, sample=f"rooted {sample_rows}-row subtree ({sample_rows // ls} leaves) of the workload, "
                                    f"nrhs {k}, numpy/OpenBLAS restatement of matmul.jl:18-62 with its per-level slice copies, "
                                    f"mean of {max(steps, 1)} runs")


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    rows = min(cfg["n"], 2 ** 17)
    s = cpu_sample(cfg, args.steps, args.warmup, rows)
    line = {
        "impl": "reference", "metric": "HSS matmul GFLOP/s", "value": s["gflops"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s["ms"],
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "note": "Julia is not installed in the image: the reference arm is the "
                   "oracle restatement of src/matmul.jl:18-62 (numpy/OpenBLAS, one dgemm per node)"},
        "cpu_baseline": {"value": s["gflops"], "unit": "GFLOP/s", "cores": s["cores"], "kind": "port", "sample": s["sample"]},
        "e2e": {"value": s["gflops"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--variant", default="default", help="comma-separated switches: generic, nograph, nccl, adjoint (time Y = A' X), nosolve (skip the ULV solver extra), ulvfast (experimental solve plan on the fixed-shape kernels)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = dict(CONFIGS[args.config])
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, cfg, rank, world)

    import torch
    import torch.distributed as dist
    import hssb200 as hb

    if not torch.cuda.is_available() or hb.device_count() < 1:
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = world
    ls, r, k = cfg["leafsize"], cfg["rank"], cfg["nrhs"]
    n_total = cfg["n"] * N if cfg["scaling"] == "weak" else cfg["n"]

    P = hb.synthetic(n_total, ls, r, SEED, device=local, shard_rank=rank, n_shards=N)
    variants = set(args.variant.split(","))
    exchange = None
    if N > 1 and "nccl" not in variants:   # default: NVLink peer stores into IPC-mapped workspaces
        ok = 1
        try:
            P.reserve(cfg["nrhs"])
            blob = P.xchg_export()
        except hb.HssbError:
            ok, blob = 0, b""
        blobs = [None] * N
        dist.all_gather_object(blobs, blob)
        if ok and all(len(b) == 128 for b in blobs):
            try:
                P.xchg_import(blobs)
            except hb.HssbError:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 1:
            exchange = "nvlink peer stores (CUDA IPC)"
        else:                              # no peer access on this box: rebuild and use NCCL
            P.close()
            P = hb.synthetic(n_total, ls, r, SEED, device=local, shard_rank=rank, n_shards=N)
    if N > 1 and exchange is None:         # exchange = one ncclAllGather per product
        uid = [hb.PackedHss.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        P.comm_init(uid[0], rank, N)
        exchange = "nccl all-gather"
    if "generic" in variants:
        P.set_option(hb.OPT_FORCE_GENERIC, 1)
    P.set_option(hb.OPT_USE_GRAPH, 0 if "nograph" in variants else 1)
    rows = P.info.local_n
    # a dedicated non-default stream: the library launches on it and torch's events time it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream
    X = torch.empty((k, rows), dtype=torch.float64, device="cuda")  # column-major rows x k
    Y = torch.empty((k, rows), dtype=torch.float64, device="cuda")
    hb._check(hb.lib().hssb_synthetic_rhs(SEED, n_total, k, P.info.local_col0, rows, X.data_ptr(), rows, local, st))
    P.reserve(k)
    flops_local, bytes_local = P.flops(k), P.algorithmic_bytes(k)

    def barrier():
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    adjoint = "adjoint" in variants  # time Y = A' X (hssb_matmul_t_dev) instead of Y = A X

    def step():
        P.matmul_dev(X.data_ptr(), rows, Y.data_ptr(), rows, k, 1.0, 0.0, stream=st, trans=adjoint)

    # ---------------- device-resident throughput (`value`) -------------------
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = P.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    tw1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = P.launch_count() - l0
    clocks = sampler.stop(tw0, tw1) if sampler else None
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(flops_local), float(bytes_local), float(launches)], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step = tmax.item() / args.steps
    flops_all, bytes_all, launches_all = tot.tolist()
    gflops = flops_all / (ms_step * 1e-3) * 1e-9
    gbs = bytes_all / (ms_step * 1e-3) * 1e-9

    # ---------------- per-kernel timing for the roofline ----------------------
    # Dominant kernel = the leaf-down kernel (Y = D X + U F: 2mk(m+r) of the 2mk(m+2r)+... flops).
    prof = None
    if rank == 0 or N > 1:
        prof = profile_phases(hb, P, X, Y, rows, k, st, max(3, min(args.steps, 10)), adjoint)

    # ---------------- end to end through the host entry (`e2e`) ---------------
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((k, rows), dtype=torch.float64).pin_memory()
        Yh = torch.empty((k, rows), dtype=torch.float64).pin_memory()
        Xh.copy_(X)
        xs, ys = Xh.numpy().T, Yh.numpy().T  # column-major (rows x k) views of the pinned buffers
        steps_e = max(3, min(args.steps, 5))
        for _ in range(2):
            P.mul_(ys, xs, 1.0, 0.0, trans=adjoint)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            P.mul_(ys, xs, 1.0, 0.0, trans=adjoint)  # H2D of X, product, D2H of Y, synchronous
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / steps_e], dtype=torch.float64, device="cuda")
        if N > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        chk = float(torch.linalg.norm(torch.from_numpy(ys[:, 0]) - Y[0].cpu()) / torch.linalg.norm(Y[0].cpu()))
        e2e = {"value": flops_all / te.item() * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * rows * k,
               "d2h_bytes_per_step": 8 * rows * k, "ms_per_step": te.item() * 1e3, "steps": steps_e,
               "host_memory": "pinned", "matches_device_path": chk <= 1e-12}

    # ---------------- the ULV solver on the same matrix (extra, config 3 only) -----
    # Not part of `value`: SURVEY 8f rank 4 measured beside the product (factor once, then solves).
    solve = None
    if N == 1 and args.config == "c3" and not adjoint and "nosolve" not in variants:
        try:
            if "ulvfast" in variants:   # experimental: fixed-shape kernels on the solve plan (HSSB_OPT_ULV_FAST)
                P.set_option(hb.OPT_ULV_FAST, 1)
            ui = P.ulv_info
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            P.ulv_factor()
            torch.cuda.synchronize()
            t_factor = time.perf_counter() - t0
            Zs = torch.empty_like(X)
            for _ in range(3):
                P.solve_dev(X.data_ptr(), rows, Zs.data_ptr(), rows, k, stream=st)
            steps_s = max(3, min(args.steps, 10))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(steps_s):
                P.solve_dev(X.data_ptr(), rows, Zs.data_ptr(), rows, k, stream=st)
            s1.record()
            torch.cuda.synchronize()
            ms_solve = s0.elapsed_time(s1) / steps_s
            P.matmul_dev(Zs.data_ptr(), rows, Y.data_ptr(), rows, k, 1.0, 0.0, stream=st)   # A (A \ X) - X
            torch.cuda.synchronize()
            resid = float(torch.linalg.norm(Y - X) / torch.linalg.norm(X))
            fl_s = ui.flops_per_rhs * k
            by_s = ui.flops_per_rhs // 2 * 8 + 2 * 8 * rows * k
            solve = {"what": "Z = A \\ X (hssb_solve_dev, ULV factors resident, device time per call)", "ms_per_solve": ms_solve,
                     "gflops": fl_s / ms_solve * 1e-6, "flops": fl_s, "algorithmic_bytes": by_s, "hbm_gbs": by_s / ms_solve * 1e-6,
                     "factor_ms_once": t_factor * 1e3, "factor_pool_gb": ui.pool_bytes * 1e-9,
                     "fast_form": P.get_option(hb.OPT_ULV_FAST) == 2,
                     "relative_residual": resid, "steps": steps_s,
                     "note": "||A Z - X|| / ||X|| with A Z from the product path; the synthetic matrix is ill conditioned "
                             "(cond ~ 1e7), the scale-free backward error ||A Z - X|| / (||A||_2 ||Z||) is asserted <= 1e-12 in "
                             "tests/test_gpu_ulv.py (measured 6e-17)"}
        except Exception as e:   # the extra must never take the bench line down
            solve = {"error": repr(e)}

    # ---------------- measured peaks + roofline --------------------------------
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        dmma = hb.measure_peak(1, 20000, local)
        dfma = hb.measure_peak(0, 20000, local)
        fp64_peak = max(dmma, dfma)
        t_mem = bytes_all / N / (hbm_peak * 1e9)
        t_flop = flops_all / N / (fp64_peak * 1e12)
        bound = "tensor" if t_flop >= t_mem else "hbm"
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.config, {}).get("leaf_down")
        except Exception:
            pass
        dom = prof["dominant"]
        roofline = {
            "kernel": dom["name"], "bound": "tensor", "achieved": dom["tflops"], "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": dom["tflops"] / fp64_peak, "traffic": traffic,
            "launch_ms": dom["ms"], "flops_per_launch": dom["flops"], "algorithmic_bytes_per_launch": dom["bytes"],
            "hbm_gbs_achieved": dom["bytes"] / (dom["ms"] * 1e-3) * 1e-9, "hbm_frac": dom["bytes"] / (dom["ms"] * 1e-3) * 1e-9 / hbm_peak,
            "peak_source": "FP64 peak measured live by hssb_measure_peak (register-resident DMMA m8n8k4 / DFMA loops; "
                           "MEASURED_PEAKS.json carries no FP64 figure); tensor = FP64 DMMA, the only tensor path for f64 on sm_100a",
            "share_of_step": dom["ms"] / prof["total_ms"],
        }
        out = {
            "metric": "HSS matmul GFLOP/s", "value": gflops, "unit": "GFLOP/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["desc"], "n_total": n_total, "leafsize": ls, "rank": r, "nrhs": k,
                       "parallelism": f"subtree-shard x{N}" if N > 1 else "single GPU", "variant": args.variant,
                       "operator": "A' (adjoint twin pool)" if adjoint and P.get_option(hb.OPT_ADJOINT_TWIN) == 2 else ("A'" if adjoint else "A"),
                       "exchange": exchange, "host_numa_node": numa,
                       "l2": "working set (generators + X + Y = %.2f GB per GPU) >> 126 MB L2, no flush needed" % (bytes_local * 1e-9),
                       "cuda_graph": bool("nograph" not in variants)},
            "hbm_gbs": gbs,
            "product_roofline": {
                "flops": flops_all, "algorithmic_bytes": bytes_all, "t_mem_ms": t_mem * 1e3, "t_flop_ms": t_flop * 1e3,
                "binding": bound, "frac_hbm": t_mem * 1e3 / ms_step, "frac_fp64": t_flop * 1e3 / ms_step,
                "frac_of_roofline": max(t_mem, t_flop) * 1e3 / ms_step,
                "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src, "fp64_dmma_tflops": dmma, "fp64_dfma_tflops": dfma},
            "roofline": roofline,
            "phases_ms": prof["phases"],
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "e2e": e2e,
            "ulv_solve": solve,
        }
    # ---------------- CPU baseline (rank 0, N = 1 only) -------------------------
    if rank == 0 and N == 1 and not args.no_cpu and isinstance(solve, dict) and "ms_per_solve" in solve:
        # the CPU restatement of ulvfactsolve (factorises + solves per call, like ulvfactor.jl) on a 2^15-row subtree
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import numpy as np
            import hss_oracle as o
            import hss_ulv_oracle as uo
            nc = min(cfg["n"], 2 ** 15)
            hc = o.synthetic_hss(nc, ls, r, SEED)
            Bc = o.synth_x(SEED, nc, k)
            tc0 = time.perf_counter()
            uo.ulvfactsolve(hc, Bc)
            tc = time.perf_counter() - tc0
            solve["cpu_restatement"] = {"seconds_per_call_sample": tc, "sample_rows": nc,
                                        "seconds_per_call_scaled": tc * cfg["n"] / nc, "kind": "port",
                                        "note": "numpy restatement of ulvfactor.jl:10-107, scaled by the number of leaves"}
        except Exception as e:   # the extra must never take the bench line down
            solve["cpu_restatement"] = {"error": repr(e)}
    if rank == 0 and N == 1 and not args.no_cpu:
        s = cpu_sample(cfg, 3, 1, min(cfg["n"], 2 ** 17))
        out["cpu_baseline"] = {"value": s["gflops"], "unit": "GFLOP/s", "cores": s["cores"], "kind": "port",
                               "sample": s["sample"], "best": s["best_gflops"]}
    P.close()
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank to the CPU cores of its GPU's NUMA node before any pinned host buffer is
    allocated (first touch), so that the end-to-end leg does not cross sockets: with one rank per
    GPU all eight PCIe links are busy at once and remote host memory becomes the bottleneck."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def profile_phases(hb, P, X, Y, rows, k, st, reps, trans=False):
    """Per-phase device times from CUDA events recorded by the library between
    its own launches on the launching stream (HSSB_OPT_PROFILE)."""
    import torch
    P.set_option(hb.OPT_PROFILE, 1)
    saved_graph = P.get_option(hb.OPT_USE_GRAPH)
    P.set_option(hb.OPT_USE_GRAPH, 0)
    acc = None
    for i in range(reps + 1):
        P.matmul_dev(X.data_ptr(), rows, Y.data_ptr(), rows, k, 1.0, 0.0, stream=st, trans=trans)
        torch.cuda.synchronize()
        ph = P.phase_times()
        if i == 0:
            continue  # warm-up of the un-graphed path
        if acc is None:
            acc = [dict(p) for p in ph]
            for a in acc:
                a["ms"] = 0.0
        for a, p in zip(acc, ph):
            a["ms"] += p["ms"] / reps
    for a in acc:
        a["flops"] = a["flops_per_rhs"] * k
        a["bytes"] = 8 * (a["gen_elems"] + (a["x_rows"] + a["y_rows"]) * k)
    P.set_option(hb.OPT_PROFILE, 0)
    P.set_option(hb.OPT_USE_GRAPH, saved_graph)
    total = sum(a["ms"] for a in acc)
    for a in acc:
        a["tflops"] = a["flops"] / (a["ms"] * 1e-3) * 1e-12 if a["ms"] > 0 else 0.0
    dom = max((a for a in acc if a["kind"] == 4), key=lambda a: a["ms"], default=max(acc, key=lambda a: a["ms"]))
    return {"phases": [{"name": a["name"], "ms": round(a["ms"], 5), "tasks": a["ntasks"], "tflops": round(a["tflops"], 3),
                        "fast": a["fast"]} for a in acc], "total_ms": total, "dominant": dom}


if __name__ == "__main__":
    main()
