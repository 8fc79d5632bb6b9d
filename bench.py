#!/usr/bin/env python
"""bench.py — HSS x dense product throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config c3|c4|c5] [--impl reference]

A step = one `hssA * X` (reference src/matmul.jl:13-62) over one synthetic right-hand side.

Headline line (`value`): BASELINE config 3 (n = 2^20 per GPU, leafsize 128, rank 32, nrhs 64 — the
configuration the north-star target is quoted on).  N > 1 (torchrun, one rank per GPU) shards the tree
by subtree: weak scaling, every GPU owns a config-3-sized subtree (n = N * 2^20), one exchange of the
subtree-root Z blocks per product.

`extra`: the multi-GPU configurations BASELINE.json names, measured in the same run with the same
protocol — config 4 (n = 2^22, rank 64, nrhs 128, STRONG scaling over the N GPUs of this run) at every N,
config 5 (n = 2^24, leafsize 256, rank 64, nrhs 32, strong) at N = 1 and N = 8.

Every measured configuration carries `parity_rel_err`: one extra product on a sparse-support right-hand
side that straddles the middle shard cut, with sampled leaves of EVERY shard compared against the lazily
evaluated oracle (oracle.LazySyntheticHss, the checker only; bar 1e-12).

Rank 0 prints ONE JSON line.
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

CONFIGS = {  # name: n per GPU (weak) or total (strong), leafsize, rank, nrhs
    "c3": dict(n=2 ** 20, leafsize=128, rank=32, nrhs=64, scaling="weak",
               desc="synthetic random-generator HSS n=2^20 per GPU, leafsize 128, rank 32, nrhs 64"),
    "c4": dict(n=2 ** 22, leafsize=128, rank=64, nrhs=128, scaling="strong",
               desc="synthetic random-generator HSS n=2^22, leafsize 128, rank 64, nrhs 128, subtree-sharded"),
    "c5": dict(n=2 ** 24, leafsize=256, rank=64, nrhs=32, scaling="strong",
               desc="synthetic random-generator HSS n=2^24, leafsize 256, rank 64, nrhs 32, subtree-sharded"),
}
SEED = 3
PARITY_TOL = 1e-12


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis and all(t.strip().isdigit() for t in vis.split(",")):
        return int(vis.split(",")[local])
    return local


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
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(gpu_index))
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
# one dgemm per node like src/matmul.jl).  Julia is not installed in the image.
# ---------------------------------------------------------------------------
def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return os.cpu_count() or 1


class blas_limit:
    """threadpoolctl limit that also RAISES the thread count (torchrun exports OMP_NUM_THREADS=1)."""

    def __init__(self, n):
        self.n, self.ctx = n, None

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
            self.ctx = threadpool_limits(limits=self.n, user_api="blas")
            self.ctx.__enter__()
        except Exception:
            self.ctx = None
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_time_product(o, np, h, X, steps, warmup, copy_slices):
    C = np.empty((X.shape[0], X.shape[1]))
    for _ in range(max(warmup, 1)):
        o.mul(C, h, X, 1.0, 0.0, copy_slices=copy_slices)
    ts = []
    for _ in range(max(steps, 1)):
        t = time.perf_counter()
        o.mul(C, h, X, 1.0, 0.0, copy_slices=copy_slices)  # copy_slices=True: the per-level copies of matmul.jl:37-38
        ts.append(time.perf_counter() - t)
    return ts


def cpu_sample(cfg, steps, warmup, sample_rows, threads=None, copy_slices=True):
    """A depth-d subtree of the workload treated as root (rooted(), matmul.jl:24): same leaves, same
    per-node GEMM shapes, 1/2^d of the work; sample_rows = n is the workload in full."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import hss_oracle as o
    ls, r, k = cfg["leafsize"], cfg["rank"], cfg["nrhs"]
    threads = threads or host_cores()
    h = o.synthetic_hss(sample_rows, ls, r, SEED)
    X = o.synth_x(SEED, sample_rows, k)
    _, flops = o.algorithmic_counts(h, k)
    with blas_limit(threads):
        ts = cpu_time_product(o, np, h, X, steps, warmup, copy_slices)
        used = blas_threads()
    best, mean = min(ts), sum(ts) / len(ts)
    what = "the workload in full" if sample_rows >= cfg["n"] else f"rooted {sample_rows}-row subtree ({sample_rows // ls} leaves) of the workload"
    return dict(gflops=flops / mean * 1e-9, best_gflops=flops / best * 1e-9, ms=mean * 1e3, cores=used, flops=flops,
                tree=(h, X), sample=f"{what}, nrhs {k}, numpy/OpenBLAS restatement of matmul.jl:18-62 "
                                    f"{'with its per-level slice copies (:37-38)' if copy_slices else 'with views instead of slice copies'}, "
                                    f"{used} BLAS threads, mean of {max(steps, 1)} runs")


def workload_config(cfg, n_total, N):
    """The keys both arms print identically."""
    return {"workload": cfg["desc"], "n_total": n_total, "leafsize": cfg["leafsize"], "rank": cfg["rank"], "nrhs": cfg["nrhs"],
            "parallelism": f"subtree-shard x{N}" if N > 1 else "single GPU"}


def run_reference(args, cfg, rank, world):
    """`--impl reference`: the reference's CPU algorithm on the box's host cores.  Julia is absent, so
    this is the oracle restatement (kind = "port").  The workload is run IN FULL when (warmup + steps)
    products fit a ~3.5 minute budget (config 3: ~1.5 s per product on 16 threads), otherwise on the
    largest subtree that does; under torchrun rank 0 runs one GPU's share of the weak-scaled workload."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import hss_oracle as o
    cores = host_cores()
    probe_rows = min(cfg["n"], 2 ** 16)
    probe = cpu_sample(cfg, 1, 1, probe_rows, threads=cores)
    probe.pop("tree")
    per_row = probe["ms"] * 1e-3 / probe_rows
    budget = 200.0
    rows = cfg["n"]
    while rows > probe_rows and (args.steps + max(args.warmup, 1) + 9) * per_row * rows > budget:
        rows //= 2
    s = cpu_sample(cfg, args.steps, args.warmup, rows, threads=cores, copy_slices=True)
    h, X = s.pop("tree")
    # both slice-copy and view variants, BLAS threads in {1, all}; bounded repetitions
    variants = {"copies_all_threads": {"gflops": s["gflops"], "ms": s["ms"], "threads": s["cores"]}}
    for name, thr, cs in (("views_all_threads", cores, False), ("copies_1_thread", 1, True), ("views_1_thread", 1, False)):
        with blas_limit(thr):
            ts = cpu_time_product(o, np, h, X, 2, 1, cs)
            variants[name] = {"gflops": s["flops"] / (sum(ts) / len(ts)) * 1e-9, "ms": sum(ts) / len(ts) * 1e3, "threads": blas_threads()}
    n_total = cfg["n"] * world if cfg["scaling"] == "weak" else cfg["n"]
    line = {
        "impl": "reference", "metric": "HSS matmul GFLOP/s", "value": s["gflops"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s["ms"],
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, n_total, world),
        "reference_note": "Julia is not installed in the image: the reference arm is the oracle restatement of "
                          "src/matmul.jl:18-62 (numpy/OpenBLAS, one dgemm per node, per-level slice copies as :37-38); "
                          "GFLOP/s is a rate, measured on `cpu_baseline.sample`",
        "cpu_baseline": {"value": s["gflops"], "unit": "GFLOP/s", "cores": s["cores"], "kind": "port", "sample": s["sample"],
                         "sample_rows": rows, "host_cores": cores, "variants": variants},
        "e2e": {"value": s["gflops"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def barrier(ctx):
    torch, dist = ctx["torch"], ctx["dist"]
    torch.cuda.synchronize()
    if ctx["N"] > 1:
        dist.barrier()
    torch.cuda.synchronize()


class Run:
    """One configuration on the N GPUs of this job: handle, device-resident X / Y, timing, parity."""

    def __init__(self, ctx, name, variants):
        self.ctx, self.name, self.cfg = ctx, name, dict(CONFIGS[name])
        hb, torch, dist = ctx["hb"], ctx["torch"], ctx["dist"]
        N, rank, local = ctx["N"], ctx["rank"], ctx["local"]
        cfg = self.cfg
        self.ls, self.r, self.k = cfg["leafsize"], cfg["rank"], cfg["nrhs"]
        self.n_total = cfg["n"] * N if cfg["scaling"] == "weak" else cfg["n"]

        def make():
            return hb.synthetic(self.n_total, self.ls, self.r, SEED, device=local, shard_rank=rank, n_shards=N)

        P = make()
        self.exchange = None
        if N > 1 and "nccl" not in variants:   # default: NVLink peer stores into IPC-mapped workspaces
            ok = 1
            try:
                P.reserve(self.k)
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
                self.exchange = "nvlink peer stores (CUDA IPC)" + (", inside the tree kernel" if "tree" in variants else ", one push kernel per product")
            else:                              # no peer access on this box: rebuild and use NCCL
                P.close()
                P = make()
        if N > 1 and self.exchange is None:    # exchange = one ncclAllGather per product
            uid = [hb.PackedHss.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            P.comm_init(uid[0], rank, N)
            self.exchange = "nccl all-gather"
        if "generic" in variants:
            P.set_option(hb.OPT_FORCE_GENERIC, 1)
        P.set_option(hb.OPT_USE_GRAPH, 0 if "nograph" in variants else 1)
        if "tree" in variants:       # all merge / translate levels in one persistent cooperative launch (opt-in)
            P.set_option(hb.OPT_TREE_KERNEL, 1)
        self.P = P
        self.rows = P.info.local_n
        self.row0 = P.info.local_col0
        self.X = torch.empty((self.k, self.rows), dtype=torch.float64, device="cuda")  # column-major rows x k
        self.Y = torch.empty((self.k, self.rows), dtype=torch.float64, device="cuda")
        hb._check(hb.lib().hssb_synthetic_rhs(SEED, self.n_total, self.k, self.row0, self.rows, self.X.data_ptr(), self.rows, local, ctx["st"]))
        P.reserve(self.k)
        self.flops_local, self.bytes_local = P.flops(self.k), P.algorithmic_bytes(self.k)
        self.adjoint = "adjoint" in variants  # time Y = A' X (hssb_matmul_t_dev) instead of Y = A X

    def step(self):
        self.P.matmul_dev(self.X.data_ptr(), self.rows, self.Y.data_ptr(), self.rows, self.k, 1.0, 0.0, stream=self.ctx["st"], trans=self.adjoint)

    def timed(self, steps, warmup, sample_clocks=False):
        torch, dist, N = self.ctx["torch"], self.ctx["dist"], self.ctx["N"]
        sampler = ClockSampler(self.ctx["local"]) if sample_clocks else None
        for _ in range(warmup):
            self.step()
        barrier(self.ctx)
        l0 = self.P.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        barrier(self.ctx)
        tw1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        launches = self.P.launch_count() - l0
        clocks = sampler.stop(tw0, tw1) if sampler else None
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        tot = torch.tensor([float(self.flops_local), float(self.bytes_local), float(launches)], dtype=torch.float64, device="cuda")
        if N > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)   # max over ranks of the device time
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        self.ms_step = tmax.item() / steps
        self.flops_all, self.bytes_all, self.launches_all = tot.tolist()
        self.gflops = self.flops_all / (self.ms_step * 1e-3) * 1e-9
        self.gbs = self.bytes_all / (self.ms_step * 1e-3) * 1e-9
        return clocks

    def parity(self):
        """Sampled-leaf check against the lazily evaluated oracle on a right-hand side supported on
        three leaves around the MIDDLE of the matrix (it straddles the cut between shards N/2-1 and N/2,
        so every sampled Y row of the other shards arrives through the exchange and the top tree).
        Returns (max over ranks of ||Y_gpu - Y_oracle||_F / ||Y_oracle||_F over the rank's sampled
        leaves, total number of leaves checked)."""
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import numpy as np
        import hss_oracle as o
        torch, dist, N = self.ctx["torch"], self.ctx["dist"], self.ctx["N"]
        n, ls, k = self.n_total, self.ls, self.k
        s_lo, s_len = n // 2 - ls - 17, 3 * ls
        Xs = np.random.default_rng(0).standard_normal((s_len, k))
        Xd = torch.zeros_like(self.X)
        a, b = max(s_lo, self.row0), min(s_lo + s_len, self.row0 + self.rows)
        if a < b:
            Xd[:, a - self.row0:b - self.row0] = torch.from_numpy(np.ascontiguousarray(Xs[a - s_lo:b - s_lo].T)).cuda()
        Yd = torch.full_like(self.Y, float("nan"))
        self.P.matmul_dev(Xd.data_ptr(), self.rows, Yd.data_ptr(), self.rows, k, 1.0, 0.0, stream=self.ctx["st"], trans=False)
        torch.cuda.synchronize()
        nl = self.rows // ls
        targets = {self.row0 + i * ls for i in (0, nl // 3, nl // 2, nl - 1)}
        for t in range(s_lo // ls * ls, s_lo + s_len, ls):   # the support leaves themselves (the D X term)
            if self.row0 <= t < self.row0 + self.rows:
                targets.add(t)
        lazy = o.LazySyntheticHss(n, ls, self.r, SEED).rows(sorted(targets), s_lo, Xs)
        num = den = 0.0
        for lo, yref in lazy.items():
            got = Yd[:, lo - self.row0:lo - self.row0 + ls].cpu().numpy().T
            d = got - yref
            num += float((d * d).sum()) if np.isfinite(got).all() else float("inf")
            den += float((yref * yref).sum())
        err = (num / max(den, 1e-300)) ** 0.5
        te = torch.tensor([err, float(len(lazy))], dtype=torch.float64, device="cuda")
        if N > 1:
            tm = te.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(te, op=dist.ReduceOp.SUM)
            return tm[0].item(), int(te[1].item())
        return te[0].item(), int(te[1].item())

    def roofline(self, peaks):
        N = self.ctx["N"]
        t_mem = self.bytes_all / N / (peaks["hbm"] * 1e9)
        t_flop = self.flops_all / N / (peaks["fp64"] * 1e12)
        return {"flops": self.flops_all, "algorithmic_bytes": self.bytes_all, "t_mem_ms": t_mem * 1e3, "t_flop_ms": t_flop * 1e3,
                "binding": "tensor" if t_flop >= t_mem else "hbm", "frac_hbm": t_mem * 1e3 / self.ms_step,
                "frac_fp64": t_flop * 1e3 / self.ms_step, "frac_of_roofline": max(t_mem, t_flop) * 1e3 / self.ms_step,
                "hbm_peak_gbs": peaks["hbm"], "hbm_peak_source": peaks["hbm_src"], "fp64_dmma_tflops": peaks["dmma"],
                "fp64_dfma_tflops": peaks["dfma"],
                "fp64_peak_source": "measured live by hssb_measure_peak (register-resident DMMA m8n8k4 / DFMA loops); "
                                    "MEASURED_PEAKS.json carries no FP64 figure"}

    def close(self):
        self.P.close()
        self.X = self.Y = None


def measured_peaks(hb, local):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    dmma = hb.measure_peak(1, 20000, local)
    dfma = hb.measure_peak(0, 20000, local)
    return {"hbm": peaks.get("hbm_gbs", 6650.0),
            "hbm_src": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
            "dmma": dmma, "dfma": dfma, "fp64": max(dmma, dfma)}


def run_extra(ctx, name, steps, warmup, peaks):
    """One of BASELINE.json's multi-GPU configurations, strong-scaled over the N GPUs of this run."""
    torch = ctx["torch"]
    try:
        run = Run(ctx, name, ctx["variants"])
        run.timed(steps, warmup)
        err, nleaf = run.parity()
        out = None
        if ctx["rank"] == 0:
            out = {"workload": run.cfg["desc"], "scaling": "strong", "n_gpus": ctx["N"], "value": run.gflops, "unit": "GFLOP/s",
                   "hbm_gbs": run.gbs, "ms_per_step": run.ms_step, "steps": steps, "warmup": warmup,
                   "gpu_launches": int(run.launches_all), "exchange": run.exchange,
                   "parity_rel_err": err, "parity_leaves_checked": nleaf, "parity_ok": err <= PARITY_TOL,
                   "product_roofline": run.roofline(peaks)}
        run.close()
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # an extra must never take the headline line down
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        return {"error": repr(e)} if ctx["rank"] == 0 else None


def run_small_tree_extra(ctx, what, n, leaf, rmin, rmax, k, steps, peaks):
    """BASELINE configs 1-2 are matrices out of a compression: ragged leaves, per-node ranks (9-20), far smaller
    than L2.  What the product costs depends only on the tree's shapes, so this times a tree of THOSE shapes
    (clustertree.jl:27-35 bisection, ranks drawn per node, random generators, registered node by node through the
    builder like any caller's HssMatrix) -- the Cauchy matrices themselves need the compression restatement,
    which lives in oracle/ and is only used by the tests (tests/test_gpu_parity.py runs them against the
    oracle).  L2 is flushed between timed products (the whole problem fits in it).  The check printed is the
    dataflow kernel against the level-by-level launches of the same plan (identical tiles: bit for bit)."""
    hb, torch = ctx["hb"], ctx["torch"]
    try:
        import numpy as np
        rng = np.random.default_rng(SEED)

        def build(lo, hi, isroot):
            m = hi - lo
            if m <= leaf:
                kk = 0 if isroot else int(rng.integers(rmin, rmax + 1))
                return hb.HssMatrix.leaf(rng.standard_normal((m, m)), rng.standard_normal((m, kk)), rng.standard_normal((m, kk)), rootnode=isroot)
            mid = lo + (m + 1) // 2
            a, b = build(lo, mid, False), build(mid, hi, False)
            (kr1, kw1), (kr2, kw2) = hb.gensize(a), hb.gensize(b)
            kk = 0 if isroot else int(rng.integers(rmin, rmax + 1))
            sc = 1.0 / np.sqrt(2.0 * max(kk, 1))
            return hb.HssMatrix.branch(a, b, rng.standard_normal((kr1, kw2)), rng.standard_normal((kr2, kw1)),
                                       sc * rng.standard_normal((kr1, kk)), sc * rng.standard_normal((kw1, kk)),
                                       sc * rng.standard_normal((kr2, kk)), sc * rng.standard_normal((kw2, kk)), rootnode=isroot)

        P = hb.pack(build(0, n, True), device=ctx["local"])
        st = ctx["st"]
        X = torch.randn((k, n), dtype=torch.float64, device="cuda")
        Y = torch.empty_like(X)
        flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")   # 256 MB > 126 MB L2
        res = {}
        # library defaults first (what a caller gets), then the alternatives of the same plan
        for name, flow, bush in (("default", None, None), ("levels", 0, 0), ("flow", 1, 0), ("bush", 1, 1)):
            if flow is not None:
                P.set_option(hb.OPT_FLOW_KERNEL, flow)
                P.set_option(hb.OPT_BUSH_KERNEL, bush)
            P.set_option(hb.OPT_USE_GRAPH, 1)
            for _ in range(3):
                P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
            l0 = P.launch_count()
            ts = []
            for _ in range(steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                P.matmul_dev(X.data_ptr(), n, Y.data_ptr(), n, k, stream=st)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            res[name] = (ts[len(ts) // 2], (P.launch_count() - l0) // steps, Y.clone())
        fl, by = P.flops(k), P.algorithmic_bytes(k)
        ms = res["default"][0]
        t_mem, t_flop = by / (peaks["hbm"] * 1e9) * 1e3, fl / (peaks["fp64"] * 1e12) * 1e3
        default_kernel = "bush" if res["default"][1] == 3 else ("flow" if res["default"][1] == 1 else "levels")
        kernels = {"bush": "leaf-up launch, every merge / translate level as one launch over bushes of the tree (csrc/hssb_bush.cuh), leaf-down launch",
                   "flow": "persistent dataflow kernel (csrc/hssb_flow.cuh)", "levels": "one launch per level with programmatic dependent launch (the library's choice under graph replay)"}
        ref = res["levels"][2]
        out = {"workload": what, "n": n, "leafsize": leaf, "ranks": [rmin, rmax], "nrhs": k, "value": fl / ms * 1e-6, "unit": "GFLOP/s",
               "hbm_gbs": by / ms * 1e-6, "ms_per_step": ms, "steps": steps, "timing": "median of per-product CUDA-event times, L2 flushed between products",
               "launches_per_product": res["default"][1], "kernel": kernels[default_kernel] + ", CUDA-graph replay",
               "alternatives": {nm: {"ms_per_step": res[nm][0], "launches_per_product": res[nm][1],
                                     "rel_diff_to_level_launches": float((res[nm][2] - ref).norm() / ref.norm())} for nm in ("levels", "flow", "bush")},
               "product_roofline": {"flops": fl, "algorithmic_bytes": by, "t_mem_ms": t_mem, "t_flop_ms": t_flop,
                                    "binding": "tensor" if t_flop >= t_mem else "hbm", "frac_of_roofline": max(t_mem, t_flop) / ms,
                                    "note": "latency-bound: 2*depth+2 dependent levels"}}
        P.close()
        del flush
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # an extra must never take the headline line down
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 4 / config 5 strong-scaling extras")
    ap.add_argument("--variant", default="default",
                    help="comma-separated switches: generic, nograph, notree (one launch per tree level), nccl, adjoint (time Y = A' X), "
                         "nosolve (skip the ULV solver extra), ulvfast (experimental solve plan on the fixed-shape kernels)")
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
    variants = set(args.variant.split(","))
    # a dedicated non-default stream: the library launches on it and torch's events time it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    ctx = dict(hb=hb, torch=torch, dist=dist, N=N, rank=rank, local=local, st=tstream.cuda_stream, variants=variants)

    run = Run(ctx, args.config, variants)
    P, X, Y, rows, k = run.P, run.X, run.Y, run.rows, run.k
    ls, r, n_total, adjoint, st = run.ls, run.r, run.n_total, run.adjoint, ctx["st"]

    # ---------------- device-resident throughput (`value`) -------------------
    clocks = run.timed(args.steps, args.warmup, sample_clocks=(rank == 0))
    ms_step, gflops, gbs = run.ms_step, run.gflops, run.gbs
    flops_all = run.flops_all
    parity_err, parity_leaves = (None, 0) if adjoint else run.parity()

    # ---------------- per-kernel timing for the roofline ----------------------
    # Dominant kernel = the leaf-down kernel (Y = D X + U F: 2mk(m+r) of the 2mk(m+2r)+... flops).
    prof = profile_phases(hb, P, X, Y, rows, k, st, max(3, min(args.steps, 10)), adjoint)

    # ---------------- end to end through the host entry (`e2e`) ---------------
    e2e = None
    if not args.no_e2e:
        import numpy as np

        def e2e_leg(xs, ys, steps_e):
            for _ in range(2):
                P.mul_(ys, xs, 1.0, 0.0, trans=adjoint)
            barrier(ctx)
            t0 = time.perf_counter()
            for _ in range(steps_e):
                P.mul_(ys, xs, 1.0, 0.0, trans=adjoint)  # H2D of X, product, D2H of Y, synchronous
            barrier(ctx)
            te = torch.tensor([(time.perf_counter() - t0) / steps_e], dtype=torch.float64, device="cuda")
            if N > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return te.item()

        steps_e = max(3, min(args.steps, 5))
        Xh = torch.empty((k, rows), dtype=torch.float64).pin_memory()
        Yh = torch.empty((k, rows), dtype=torch.float64).pin_memory()
        Xh.copy_(X)
        # what the box's PCIe / host-memory fabric gives THIS rank while every rank of the job copies both ways at
        # once (plain cudaMemcpyAsync of the same pinned buffers, no kernels): the floor of the end-to-end leg
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dtmp = torch.empty_like(X)

        def duplex():
            with torch.cuda.stream(s_in):
                dtmp.copy_(Xh, non_blocking=True)
            with torch.cuda.stream(s_out):
                Yh.copy_(Y, non_blocking=True)
        duplex()
        torch.cuda.synchronize()
        barrier(ctx)
        t0 = time.perf_counter()
        for _ in range(3):
            duplex()
        torch.cuda.synchronize()
        barrier(ctx)
        tc = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device="cuda")
        if N > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        t_copy = tc.item()
        del dtmp
        xs, ys = Xh.numpy().T, Yh.numpy().T  # column-major (rows x k) views of the pinned buffers
        t_pin = e2e_leg(xs, ys, steps_e)
        chk = float(torch.linalg.norm(torch.from_numpy(ys[:, 0]) - Y[0].cpu()) / torch.linalg.norm(Y[0].cpu()))
        # the same call on ordinary pageable memory, as Julia's `similar(B, ...)` (matmul.jl:13) hands it over
        xp = np.array(xs, order="F", copy=True)
        yp = np.empty((rows, k), order="F")
        t_page = e2e_leg(xp, yp, steps_e)
        staged = int(P.get_option(hb.OPT_LAST_BOUNCE))
        chk_p = float(np.linalg.norm(yp[:, 0] - ys[:, 0]) / max(np.linalg.norm(ys[:, 0]), 1e-300))
        e2e = {"value": flops_all / t_pin * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * rows * k,
               "d2h_bytes_per_step": 8 * rows * k, "ms_per_step": t_pin * 1e3, "steps": steps_e,
               "host_memory": "pinned", "matches_device_path": chk <= 1e-12,
               "copy_floor": {"ms_per_step": t_copy * 1e3, "gbs_per_direction_per_gpu": 8 * rows * k / t_copy * 1e-9,
                              "what": "the same H2D and D2H bytes as plain concurrent cudaMemcpyAsync from / to the pinned buffers, all "
                                      "ranks at once, no kernels: what the box's PCIe / host-memory fabric allows at this N"},
               "pageable": {"value": flops_all / t_page * 1e-9, "unit": "GFLOP/s", "ms_per_step": t_page * 1e3,
                            "host_memory": "pageable (numpy arrays, as a Julia Matrix would be)",
                            "staging": ("library ring of pinned 2 MiB slots + worker threads (csrc/hssb_hostpipe.h)"
                                        if staged == 3 else f"driver (cudaMemcpy2DAsync on the caller's pointer), ring bits {staged}"),
                            "host_threads_per_direction": int(P.get_option(hb.OPT_HOST_THREADS)),
                            "vs_pinned": t_page / t_pin, "matches_pinned": chk_p <= 1e-12}}
        del Xh, Yh, xp, yp

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
                     "factor_ms_once": t_factor * 1e3, "factor_device_ms": P.get_option(hb.OPT_LAST_FACTOR_US) * 1e-3,
                     "factor_note": "factor_ms_once is the wall time of hssb_ulv_factor in this process (it allocates and clears the factor pool and "
                                    "its scratch beside torch's cached blocks); factor_device_ms is the device time of its kernels (CUDA events)",
                     "factor_pool_gb": ui.pool_bytes * 1e-9,
                     "fast_form": P.get_option(hb.OPT_ULV_FAST) == 2,
                     "relative_residual": resid, "steps": steps_s,
                     "note": "||A Z - X|| / ||X|| with A Z from the product path; the synthetic matrix is ill conditioned "
                             "(cond ~ 1e7), the scale-free backward error ||A Z - X|| / (||A||_2 ||Z||) is asserted <= 1e-12 in "
                             "tests/test_gpu_ulv.py (measured 6e-17)"}
        except Exception as e:   # the extra must never take the bench line down
            solve = {"error": repr(e)}

    # ---------------- measured peaks + roofline --------------------------------
    out = None
    peaks = measured_peaks(hb, local) if rank == 0 else None
    if rank == 0:
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(args.config, {}).get("leaf_down")
            traffic_src = tj.get("_source", "static: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed "
                                            "`ncu --set full` capture under profiles/, not re-measured in this run")
        except Exception:
            pass
        dom = prof["dominant"]
        t_flop_k = dom["flops"] / (peaks["fp64"] * 1e12)
        t_mem_k = dom["bytes"] / (peaks["hbm"] * 1e9)
        tensor_bound = t_flop_k >= t_mem_k   # the slower of the two bounds binds (north_star)
        achieved_gbs = dom["bytes"] / (dom["ms"] * 1e-3) * 1e-9
        roofline = {
            "kernel": dom["name"], "bound": "tensor" if tensor_bound else "hbm",
            "achieved": dom["tflops"] if tensor_bound else achieved_gbs, "peak": peaks["fp64"] if tensor_bound else peaks["hbm"],
            "unit": "TFLOP/s" if tensor_bound else "GB/s",
            "frac": dom["tflops"] / peaks["fp64"] if tensor_bound else achieved_gbs / peaks["hbm"],
            "traffic": traffic, "traffic_source": traffic_src,
            "launch_ms": dom["ms"], "flops_per_launch": dom["flops"], "algorithmic_bytes_per_launch": dom["bytes"],
            "t_flop_ms_at_peak": t_flop_k * 1e3, "t_mem_ms_at_peak": t_mem_k * 1e3,
            "hbm_gbs_achieved": achieved_gbs, "hbm_frac": achieved_gbs / peaks["hbm"],
            "peak_source": "FP64 peak measured live by hssb_measure_peak (register-resident DMMA m8n8k4 / DFMA loops; "
                           "MEASURED_PEAKS.json carries no FP64 figure); tensor = FP64 DMMA, the only tensor path for f64 on sm_100a; "
                           "HBM peak: " + peaks["hbm_src"],
            "share_of_step": dom["ms"] / prof["total_ms"],
        }
        conf = workload_config(cfg, n_total, N)
        conf.update({"variant": args.variant,
                     "operator": "A' (adjoint twin pool)" if adjoint and P.get_option(hb.OPT_ADJOINT_TWIN) == 2 else ("A'" if adjoint else "A"),
                     "exchange": run.exchange, "host_numa": numa,
                     "l2": "working set (generators + X + Y = %.2f GB per GPU) >> 126 MB L2, no flush needed" % (run.bytes_local * 1e-9),
                     "cuda_graph": bool("nograph" not in variants), "tree_kernel": P.get_option(hb.OPT_TREE_KERNEL)})
        out = {
            "metric": "HSS matmul GFLOP/s", "value": gflops, "unit": "GFLOP/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": conf,
            "hbm_gbs": gbs,
            "parity_rel_err": parity_err, "parity_leaves_checked": parity_leaves,
            "parity_ok": (parity_err is not None and parity_err <= PARITY_TOL),
            "product_roofline": run.roofline(peaks),
            "roofline": roofline,
            "phases_ms": prof["phases"],
            "tree_ms": prof["tree_ms"],
            "gpu_launches": int(run.launches_all),
            "launches_per_product": run.launches_all / N / args.steps,
            "clocks": clocks,
            "e2e": e2e,
            "ulv_solve": solve,
        }
    # ---------------- CPU baseline (rank 0, N = 1 only) -------------------------
    if rank == 0 and N == 1 and not args.no_cpu and isinstance(solve, dict) and "ms_per_solve" in solve:
        # the CPU restatement of ulvfactsolve (factorises + solves per call, like ulvfactor.jl) on a 2^15-row subtree
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
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
        s = cpu_sample(cfg, 3, 1, min(cfg["n"], 2 ** 18))   # bounded sample: ~10-20 s of CPU work
        s.pop("tree")
        out["cpu_baseline"] = {"value": s["gflops"], "unit": "GFLOP/s", "cores": s["cores"], "kind": "port",
                               "sample": s["sample"], "best": s["best_gflops"]}
    run.close()
    torch.cuda.empty_cache()

    # ---------------- BASELINE's multi-GPU configurations, same run ---------------
    if args.config == "c3" and not args.no_extra and not adjoint:
        extra = {}
        for nm in ["c4"] + (["c5"] if N in (1, 8) else []):
            res = run_extra(ctx, nm, max(3, min(args.steps, 10)), 3, peaks)
            if rank == 0:
                extra[nm] = res
        if rank == 0 and N == 1:   # shapes of BASELINE configs 1-2 (variable-rank trees, any-shape kernels)
            extra["c1_shape"] = run_small_tree_extra(ctx, "tree of config 1's shape: bisection of n=2001 with leafsize 64 (62/63-row leaves), per-node ranks 9-20, nrhs 16",
                                                     2001, 64, 9, 20, 16, 20, peaks)
            extra["c2_shape"] = run_small_tree_extra(ctx, "tree of config 2's shape: n=2^16, leafsize 64, per-node ranks 13-20, nrhs 64",
                                                     2 ** 16, 64, 13, 20, 64, 20, peaks)
        if rank == 0:
            out["extra"] = extra
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank to the CPU cores / memory node of its GPU before any pinned host buffer is
    allocated (first touch), so that the end-to-end leg does not cross sockets.  Sources, in order:
    sysfs numa_node of the PCI device, NVML's CPU affinity of the device (what `nvidia-smi topo -m`
    prints).  Returns what was found and done (reported in config.host_numa)."""
    info = {"node": None, "source": None, "cpus_bound": None}
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        info["numa_nodes_visible"] = len(nodes)
    except Exception:
        pass
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info.update(node=node, source="sysfs", cpus_bound=len(cpus))
                return info
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(local))
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        if cpus and (cpus & allowed) and (cpus & allowed) != allowed:
            os.sched_setaffinity(0, cpus & allowed)
            info.update(source="nvml cpu affinity", cpus_bound=len(cpus & allowed))
        else:
            info.update(source="nvml cpu affinity covers every allowed cpu (one NUMA domain visible to this process)",
                        cpus_bound=len(allowed))
    except Exception as e:
        info["source"] = "unavailable: " + repr(e)[:80]
    return info


def profile_phases(hb, P, X, Y, rows, k, st, reps, trans=False):
    """Per-phase device times from CUDA events recorded by the library between
    its own launches on the launching stream (HSSB_OPT_PROFILE).  With the persistent tree kernel all
    merge / translate levels are one launch: its time is reported on the first covered phase."""
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
        a["tflops"] = a["flops"] / (a["ms"] * 1e-3) * 1e-12 if a["ms"] > 1e-4 else 0.0
    tree = [a for a in acc if a["kind"] in (1, 2, 3, 5)]
    tree_ms = {"ms": sum(a["ms"] for a in tree), "levels": sum(1 for a in tree if a["kind"] in (1, 3)),
               "flops": sum(a["flops"] for a in tree)}
    tree_ms["tflops"] = tree_ms["flops"] / (tree_ms["ms"] * 1e-3) * 1e-12 if tree_ms["ms"] > 0 else 0.0
    dom = max((a for a in acc if a["kind"] == 4), key=lambda a: a["ms"], default=max(acc, key=lambda a: a["ms"]))
    return {"phases": [{"name": a["name"], "ms": round(a["ms"], 5), "tasks": a["ntasks"], "tflops": round(a["tflops"], 3),
                        "fast": a["fast"]} for a in acc], "total_ms": total, "dominant": dom, "tree_ms": tree_ms}


if __name__ == "__main__":
    main()
