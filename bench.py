#!/usr/bin/env python
"""Benchmark of the hot path: the implicit kernel-matrix-vector product inside the CG solve of the reduced LS-SVM system.

    python bench.py --gpus N --steps K --warmup W            # our arm (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own OpenMP kernels on the host cores

A *step* is one CG iteration (gpu_csvm.hpp:568-636): one implicit matvec Ad = Q~ d over the whole data set plus the vector
updates (every 50th iteration a second matvec recomputes the residual).  Metric = algorithmic matvec TFLOP/s inside the CG
loop: F = d * n * (n + 1) FLOPs per matvec x matvecs executed in the timed region / device time;
`cg_iters_per_s` = timed iterations / the same time.  Workload at every N: BASELINE.json configs[1],
65,536 x 4,096 dense, RBF gamma = 1/d, fp64 (strong scaling: tiles of the triangle are sharded over the ranks).

Timing: `value` — data resident in HBM, W untimed + exactly K timed iterations, device time from CUDA events recorded by
the library on its launching stream, bracketed by barrier + synchronize, max over ranks.  `e2e` — one
plssvm_b200_solve_f64 call on PINNED HOST buffers (upload of X and y, q-kernel, r0, K iterations, download of alpha).
X (2.1 GB) is larger than L2 (126 MB) and streamed in full by every iteration, so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, d, kernel, dtype, description)
    "C1": (5000, 1000, "linear", "float64", "5,000 x 1,000 dense, linear, fp64"),
    "C2": (65536, 4096, "rbf", "float64", "65,536 x 4,096 dense, RBF gamma=1/d, fp64"),
    "C3": (131072, 1024, "polynomial", "float32", "131,072 x 1,024 dense, polynomial degree 3, fp32"),
    "C4": (262144, 2048, "linear", "float64", "262,144 x 2,048 dense, linear, fp64"),
    # prediction: N = number of support vectors; test points are processed in steps of PREDICT_STEP_POINTS
    "C5": (65536, 4096, "rbf", "float64", "predict 1,048,576 test points against a 65,536-SV RBF model (d = 4,096), fp64"),
}
PREDICT_STEP_POINTS = 65536
KERNEL_IDS = {"linear": 0, "polynomial": 1, "rbf": 2}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60, help="timed CG iterations (default 60 = SURVEY.md §8d: with 3 warm-up iterations the timed region contains the iter % 50 == 49 residual refresh)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the number of data points (development only; marks the line as non-headline)")
    ap.add_argument("--features", type=int, default=0, help="override the number of features (development only)")
    ap.add_argument("--tile-impl", type=int, default=0, help="0 auto (fp64: int8 slices on tcgen05, fp32: tcgen05 3xTF32), 1 SIMT tiles, 2 fp64 DMMA / fp32 3xTF32 tiles, 6 fp64 int8-slice tiles")
    ap.add_argument("--no-dmma-line", action="store_true", help="fp64 only: skip the short extra run of the native-FP64 DMMA tiles reported under 'fp64_dmma_tiles'")
    ap.add_argument("--linear-factorized", action="store_true", help="linear kernel only: time the factorised X (X^T v) matvec (HBM-bound) instead of the implicit tiles")
    ap.add_argument("--full-solve", action="store_true", help="additionally run the whole fit to eps = 1e-8 (fp64) / 1e-4 (fp32) through the C ABI; with C1 also on the CPU reference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline sample")
    return ap.parse_args()


def matvec_flops(N: int, d: int) -> float:
    n = N - 1
    return float(d) * n * (n + 1)


# ---- clocks -----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ---- data -------------------------------------------------------------------------------------------------------------------
def make_device_data(N, d, dtype, seed, device):
    """SURVEY.md §8d data family generated on the GPU: X ~ U(-1, 1) + 0.25 y u, balanced +-1 labels in a fixed permutation."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    tdt = {"float64": torch.float64, "float32": torch.float32}[dtype]
    X = torch.empty((N, d), dtype=tdt, device=device)
    chunk = max(1, (1 << 28) // d)
    for r0 in range(0, N, chunk):  # chunked so the fp64 uniform generator never needs a second full-size temporary
        r1 = min(N, r0 + chunk)
        X[r0:r1].uniform_(-1.0, 1.0, generator=g)
    y = torch.ones(N, dtype=tdt, device=device)
    y[N // 2:] = -1.0
    y = y[torch.randperm(N, generator=g, device=device)]
    u = torch.randn(d, generator=g, device=device, dtype=tdt)
    u /= u.norm()
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk)
        X[r0:r1] += 0.25 * y[r0:r1, None] * u[None, :]
    return X, y


def make_host_data(N, d, dtype, seed):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from datagen import make_data
    return make_data(N, d, seed, np.dtype(dtype))


def cpu_matvec_sample(kind_pref, d, kernel, dtype, budget_s, threads=None):
    """Times the reference's OpenMP matvec (oracle/_ref, else the restated port) on the leading n' rows of the workload."""
    import numpy as np
    import oracle
    kind = kind_pref if oracle.available(kind_pref) else "port"
    orc = oracle.Oracle(kind)
    # all host threads this process may use — torchrun exports OMP_NUM_THREADS=1, which would make the CPU arm single-threaded
    orc.set_threads(threads or len(os.sched_getaffinity(0)))
    cores = orc.max_threads()
    kid = KERNEL_IDS[kernel]

    def run(n_rows):
        X, _ = make_host_data(n_rows + 1, d, dtype, 4242)
        q = orc.q(kid, X, gamma=1.0 / d)
        v = np.random.default_rng(1).uniform(1, 2, n_rows).astype(X.dtype)
        t0 = time.perf_counter()
        orc.matvec(kid, X, q, v, np.zeros(n_rows, X.dtype), 2.0, 1.0, 1.0, gamma=1.0 / d)
        return time.perf_counter() - t0

    t_small = run(1024)  # calibration: large enough that all threads get 64x64 blocks
    rate = d * 1024.0 * 1025.0 / max(t_small, 1e-9)
    n_rows = int(min(8192, max(512, (budget_s * rate / d) ** 0.5)))
    n_rows -= n_rows % 64
    return orc, kind, cores, n_rows, run


# ---- reference arm ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, d, kernel, dtype, desc = WORKLOADS[args.workload]
    d = args.features or d
    per_step_budget = max(1.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    orc, kind, cores, n_rows, run = cpu_matvec_sample("reference", d, kernel, dtype, per_step_budget)
    for _ in range(args.warmup):
        run(n_rows)
    times = [run(n_rows) for _ in range(args.steps)]
    total = sum(times)
    F = d * float(n_rows) * (n_rows + 1)
    value = F * args.steps / total / 1e12
    sample = f"one OpenMP matvec (the CG iteration's dominant op) per step on the leading {n_rows} rows x {d} features of the workload, all {cores} host threads"
    line = {
        "impl": "reference", "metric": "cg_matvec_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if dtype == "float64" else "f32",
        "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}", "sample_rows": n_rows, "kernel": kernel},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": "reference" if kind == "reference" else "port", "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cg_iters_per_s_extrapolated": value * 1e12 / matvec_flops(N, d), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- our arm ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import plssvm_b200 as pb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    N, d, kernel, dtype, desc = WORKLOADS[args.workload]
    headline = not (args.rows or args.features)
    N, d = args.rows or N, args.features or d
    F = matvec_flops(N, d)
    npdt = np.dtype(dtype)

    be = pb.Backend(local_rank)
    if args.tile_impl:
        be.set_option("impl", args.tile_impl)
    if args.linear_factorized:
        be.set_option("linear_factorized", 1)
    if world > 1:
        be.init_comm_from_torch()

    X, y = make_device_data(N, d, dtype, 42 + list(WORKLOADS).index(args.workload), device)
    y_host = y.cpu().numpy()
    ds = be.dataset(X)

    # ---- device-resident timed region: W warm-up + exactly K CG iterations -------------------------------------------------------
    eps = 1e-30 if dtype == "float64" else 1e-18  # never met: the iteration count is fixed
    be.set_option("ignore_convergence", 1)  # fp32 CG can hit an exactly-zero residual after ~15 iterations on this data; time exactly K iterations
    cg = be.cg_begin(ds, y_host, kernel, eps=eps)
    cg.step(args.warmup)
    t_before = be.timings()
    barrier()
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        done_iters, _ = cg.step(args.steps)
        barrier()
        wall = time.perf_counter() - wall0
    t_after = be.timings()
    res = cg.finish()
    dev_ms = t_after["cg_loop_ms"] - t_before["cg_loop_ms"]
    tile_ms = t_after["matvec_tile_ms"] - t_before["matvec_tile_ms"]
    tile_calls = t_after["matvec_calls"] - t_before["matvec_calls"]
    launches = t_after["kernel_launches"] - t_before["kernel_launches"]
    if world > 1:
        tt = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(tt[0]), float(tt[1]) / 1e3
    assert done_iters == args.warmup + args.steps, (done_iters, args.warmup, args.steps)
    # matvec throughput: every implicit matvec of the timed region counts (K iterations + the residual refresh every 50th iteration)
    value = F * tile_calls / (dev_ms * 1e-3) / 1e12

    # ---- end to end through the C ABI with pinned host buffers --------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((N, d), dtype=X.dtype, pin_memory=True)
        Xh.copy_(X)
        yh = torch.empty(N, dtype=X.dtype, pin_memory=True)
        yh.copy_(y)
        del X
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        r2 = be.solve(Xh, yh, kernel, eps=eps, max_iter=args.steps)
        barrier()
        t_e2e = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([t_e2e], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt[0])
        t2 = be.timings()
        e2e = {"value": F * t2["matvec_calls"] / t_e2e / 1e12, "unit": "TFLOP/s", "matvecs": int(t2["matvec_calls"]), "cg_iters_per_s": r2["iterations"] / t_e2e, "h2d_bytes_per_step": t2["h2d_bytes"] / r2["iterations"],
               "d2h_bytes_per_step": t2["d2h_bytes"] / r2["iterations"], "seconds": t_e2e, "iterations": r2["iterations"],
               "note": "one plssvm_b200_solve call: H2D of X and y from pinned memory + q-kernel + r0 matvec + K iterations + D2H of alpha; bytes are per call / K"}

    # ---- optional: the whole fit to the parity tolerance, GPU vs the reference's CPU path (C1 is run to convergence)
    full = None
    be.set_option("ignore_convergence", 0)
    if args.full_solve and not args.no_e2e:
        feps = 1e-8 if dtype == "float64" else 1e-4
        barrier()
        t0 = time.perf_counter()
        rf = be.solve(Xh, yh, kernel, eps=feps)
        barrier()
        full = {"eps": feps, "gpu_seconds": time.perf_counter() - t0, "gpu_iterations": rf["iterations"]}
        if rank == 0 and world == 1 and args.workload == "C1":
            import oracle
            orc = oracle.Oracle("reference" if oracle.available("reference") else "port")
            t0 = time.perf_counter()
            rc = orc.solve(KERNEL_IDS[kernel], Xh.numpy(), yh.numpy(), gamma=1.0 / d, eps=feps)
            full.update({"cpu_seconds": time.perf_counter() - t0, "cpu_iterations": rc["iterations"], "cpu_threads": orc.max_threads(), "cpu_kind": orc.reported_kind(),
                         "alpha_max_rel_diff": float(np.max(np.abs(rf["alpha"] - rc["alpha"])) / np.max(np.abs(rc["alpha"]))) if rf["iterations"] == rc["iterations"] else None})
            full["speedup"] = full["cpu_seconds"] / full["gpu_seconds"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the tile kernel) ---------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "profiles", "peaks_b200.json")
    peak, peak_src = 37.0, "vendor figure (fallback: profiles/peaks_b200.json missing)"
    impl_used = t_after["impl_used"]
    tensor = impl_used in (2, 6, 7)
    key = "dmma_tflops_sustained_3s" if dtype == "float64" else ("cublas_sgemm_tf32_random_tflops_sustained_4s" if tensor else "ffma_tflops")
    pk = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    if key in pk:
        peak, peak_src = float(pk[key]), f"measured on this pool's B200 by tools/peak_probe ({key}; profiles/peaks_b200.json)"
        if dtype != "float64" and tensor:  # 3xTF32: three TF32 MMAs per algorithmic fp32 product
            peak, peak_src = peak / 3.0, peak_src + " / 3 (3xTF32 split)"
    fp_pipe_peak = peak  # fp64: DMMA issue peak; fp32: TF32 tensor peak / 3
    i8_products = 28.0 if dtype == "float64" else (10.0 if impl_used == 7 else 6.0)
    if impl_used in (6, 7):
        # int8-slice tiles: 28 (fp64: S = 7 slices, digit diagonals p + q >= 6), 6 (fp32: S = 3) or 10 (fp32 with --tile-impl 7: S = 4) int8
        # tensor-core MACs per algorithmic MAC, so the roofline of this kernel is the int8 tensor pipe / 28 (/ 6, / 10): tcgen05.mma kind::i8 issue-loop peak with random operands, measured
        # by tools/i8_peak_probe (burst when the kernel is timed alone, the sustained figure inside a long step; profiles/r01/i8_peaks_b200.json)
        i8_path = os.path.join(ROOT, "profiles", "r01", "i8_peaks_b200.json")
        i8 = json.load(open(i8_path)) if os.path.exists(i8_path) else {}
        long_step = args.steps * (dev_ms / max(args.steps, 1)) > 2000.0
        k8 = "i8_mma_n256_random_tops_sustained_3s" if long_step else "i8_mma_n256_random_tops_burst"
        peak = float(i8.get(k8, 4500.0)) / i8_products
        peak_src = (f"int8 tensor pipe / {int(i8_products)} int8 products per {'fp64' if dtype == 'float64' else 'fp32'} product: tcgen05.mma kind::i8 issue-loop peak, random operands "
                    f"({k8} = {i8.get(k8, 'nominal 4500')} TOPS, measured on this pool's B200 by tools/i8_peak_probe; profiles/r01/i8_peaks_b200.json)")
    avg_tile_s = tile_ms / max(tile_calls, 1) * 1e-3
    achieved = (F / world) / avg_tile_s / 1e12 if avg_tile_s > 0 else 0.0
    traffic, traffic_i8 = None, None
    ncu_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(ncu_path):
        traffic = json.load(open(ncu_path)).get(args.workload)
        traffic_i8 = json.load(open(ncu_path)).get(args.workload + "_i8")
    if t_after["impl_used"] == 3:
        # factorised linear matvec: HBM-bound, 2 passes over X per matvec; report bytes/s of the whole matvec against the copy bandwidth
        mv_s = (t_after["matvec_ms"] - t_before["matvec_ms"]) / max(tile_calls, 1) * 1e-3
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        gbs = 2.0 * (N - 1) * d * npdt.itemsize / mv_s / 1e9
        print(json.dumps({"metric": "linear_factorized_matvec", "ms_per_matvec": mv_s * 1e3, "ms_per_step": dev_ms / args.steps, "cg_iters_per_s": args.steps / (dev_ms * 1e-3),
                          "equivalent_implicit_tflops": value, "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None},
                          "config": {"workload": f"{args.workload}: {desc} [factorised linear fast path, SURVEY 8f row 4]"}, "n_gpus": world, "clocks": clocks.summary()}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    i8_name = f"tile_kernel_i8 ({'fp64' if dtype == 'float64' else 'fp32'} through {7 if dtype == 'float64' else (4 if impl_used == 7 else 3)} int8 slices, tcgen05 kind::i8)"
    kname = {6: i8_name, 7: i8_name, 2: "tile_kernel_dmma" if dtype == "float64" else "tile_kernel_tf32 (tcgen05 3xTF32)"}.get(impl_used)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic if impl_used == 2 else None,
                "kernel": kname + f"<{kernel}, sym>" if tensor else "tile_kernel_simt", "peak_source": peak_src,
                "avg_launch_ms": avg_tile_s * 1e3, "launches_timed": int(tile_calls), "flops_per_launch": F / world}
    if impl_used in (6, 7):
        roofline["int8_tops"] = achieved * i8_products
        # > 1: the result is produced faster than the floating-point pipe of that precision could (fp64: DMMA = DFMA issue peak; fp32: TF32 tensor peak / 3)
        roofline["vs_float_pipe_peak"] = achieved / fp_pipe_peak
        roofline["float_pipe_peak_tflops"] = fp_pipe_peak
        if traffic_i8 is not None:
            roofline["traffic"] = traffic_i8

    # fp64: the native-FP64 DMMA tiles (north-star kernel, tile_dmma.cuh) on the same resident data, a few iterations, against the DMMA peak
    dmma_line = None
    if dtype == "float64" and impl_used == 6 and world == 1 and not args.no_dmma_line:
        be.set_option("impl", 2)
        be.set_option("ignore_convergence", 1)
        cg2 = be.cg_begin(ds, y_host, kernel, eps=eps)
        cg2.step(1)
        tb = be.timings()
        cg2.step(3)
        ta = be.timings()
        cg2.finish()
        be.set_option("ignore_convergence", 0)
        be.set_option("impl", args.tile_impl)
        ms2 = (ta["matvec_tile_ms"] - tb["matvec_tile_ms"]) / max(ta["matvec_calls"] - tb["matvec_calls"], 1)
        a2 = (F / world) / (ms2 * 1e-3) / 1e12
        dmma_line = {"kernel": f"tile_kernel_dmma<{kernel}, sym>", "avg_launch_ms": ms2, "achieved": a2, "peak": fp_pipe_peak, "unit": "TFLOP/s", "frac": a2 / fp_pipe_peak,
                     "traffic": traffic, "launches_timed": int(ta["matvec_calls"] - tb["matvec_calls"]),
                     "note": "the same matvec with --tile-impl 2: TMA + mma.sync m8n8k4.f64 tiles against the measured DMMA issue peak"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        orc, kind, cores, n_rows, run = cpu_matvec_sample("reference", d, kernel, dtype, args.cpu_seconds)
        t_cpu = run(n_rows)
        cpu_baseline = {"value": d * float(n_rows) * (n_rows + 1) / t_cpu / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "reference" if kind == "reference" else "port",
                        "sample": f"one OpenMP matvec on the leading {n_rows} rows x {d} features ({t_cpu:.1f} s); same kernel / real type"}

    # optional second baseline: the reference's own CUDA kernel on this GPU, on a slice of the workload
    ref_cuda = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            if oracle.RefCuda.available():
                rows = min(N, 8192 + 1)
                Xs, _ = make_device_data(rows, d, dtype, 4242, device)
                ds_s = be.dataset(Xs)
                q_s, k_s = be.run_q_kernel(ds_s, kernel)
                v_s = np.ones(rows - 1, npdt)
                _, ms_ref = oracle.RefCuda().matvec(KERNEL_IDS[kernel], Xs, q_s, v_s, np.zeros(rows - 1, npdt), float(k_s) + 1.0, 1.0, 1.0, gamma=1.0 / d, reps=3)
                be.run_svm_kernel(ds_s, q_s, v_s, np.zeros(rows - 1, npdt), float(k_s) + 1.0, 1.0, 1.0, kernel)
                ms_ours = be.timings()["matvec_tile_ms"]
                Fs = matvec_flops(rows, d)
                ref_cuda = {"what": "reference CUDA kernel (svm_kernel.cu, compiled unchanged for sm_100, reference grid/layout) vs ours, one matvec on the same B200",
                            "rows": rows, "features": d, "reference_ms": ms_ref, "reference_tflops": Fs / ms_ref / 1e9, "ours_ms": ms_ours, "ours_tflops": Fs / ms_ours / 1e9}
                ds_s.close()
        except Exception as e:  # the baseline is optional: never fail the benchmark because of it
            ref_cuda = {"error": str(e)[:200]}

    line = {
        "metric": "cg_matvec_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if dtype == "float64" else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}" + ("" if headline else f" [DEV OVERRIDE rows={N} features={d}]"), "kernel": kernel, "rows": N, "features": d,
                   "flops_per_step": F, "l2": "inputs (2.1 GB) larger than L2; no flush needed", "parallelism": f"triangle tiles sharded over {world} rank(s), X replicated"},
        "cg_iters_per_s": args.steps / (dev_ms * 1e-3), "wall_ms_per_step": wall / args.steps * 1e3,
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
        "final_residual": float(res["delta"]), "tile_impl": int(t_after["impl_used"]), "full_solve": full,
        "precision_note": (("fp64 storage, vectors and epilogue; x_i.x_j through 7 int8 digit planes per operand, 28 exact int32 tensor-core products recombined in fp64 "
                            "(error vs the fp64 oracle <= that of the DMMA tiles)" if impl_used == 6 else None) if dtype == "float64" else
                           ("fp32 storage, vectors and epilogue; x_i.x_j through 3 int8 digit planes per operand (22 bits relative to the row maximum = the input precision of the 3xTF32 "
                            "scheme), 6 exact int32 tensor-core products recombined in fp64, rounded once to fp32 (error vs the fp64 oracle ~2e-7: at or below the 3xTF32 tiles and "
                            "at the level of fp32 FMA tiles; --tile-impl 7 = 4 planes / 30 bits)" if impl_used == 6 else
                            "fp32 storage, vectors and epilogue; x_i.x_j through 4 int8 digit planes per operand (30 bits), 10 exact int32 tensor-core products recombined in fp64, rounded "
                            "once to fp32" if impl_used == 7 else
                            "fp32 storage and accumulation; products via the 3xTF32 split on tcgen05 (error vs fp64 oracle ~2e-7, same as FFMA fp32)")),
        "matvecs_in_timed_region": int(tile_calls), "reference_cuda_baseline": ref_cuda, "fp64_dmma_tiles": dmma_line,
    }
    print(json.dumps(_finite(line)), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- prediction workload (C5): run_predict_kernel throughput -----------------------------------------------------------------------
def run_predict(args):
    """A step = decision values of PREDICT_STEP_POINTS test points against all support vectors (2 m n_sv d FLOPs)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import plssvm_b200 as pb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    n_sv, d, kernel, dtype, desc = WORKLOADS["C5"]
    n_sv, d = args.rows or n_sv, args.features or d
    m = PREDICT_STEP_POINTS  # per rank and step: test points are independent units, sharded over ranks with no collective
    F = 2.0 * m * n_sv * d
    be = pb.Backend(local_rank)
    if args.tile_impl:
        be.set_option("impl", args.tile_impl)
    SV, _ = make_device_data(n_sv, d, dtype, 47, device)
    P, _ = make_device_data(m, d, dtype, 48 + rank, device)
    rng = np.random.default_rng(47)
    alpha = rng.uniform(-1, 1, n_sv)
    alpha -= alpha.mean()
    rho = 0.1
    sv_ds, p_ds = be.dataset(SV), be.dataset(P)
    for _ in range(args.warmup):
        be.predict_values(sv_ds, alpha, rho, p_ds, kernel)
    barrier()
    tile_ms, launches = 0.0, 0
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            vals, _ = be.predict_values(sv_ds, alpha, rho, p_ds, kernel)
            t = be.timings()
            tile_ms += t["matvec_tile_ms"]
            launches += t["kernel_launches"]
        barrier()
        wall = time.perf_counter() - t0
    e2e = None
    if not args.no_e2e:
        Ph = torch.empty((m, d), dtype=P.dtype, pin_memory=True)
        Ph.copy_(P)
        SVh = torch.empty((n_sv, d), dtype=P.dtype, pin_memory=True)
        SVh.copy_(SV)
        barrier()
        t0 = time.perf_counter()
        vals2, _ = be.predict_values(SVh, alpha, rho, Ph, kernel)
        barrier()
        t_e2e = time.perf_counter() - t0
        t2 = be.timings()
        assert np.array_equal(vals, vals2)
        e2e = {"value": world * F / t_e2e / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": t2["h2d_bytes"], "d2h_bytes_per_step": t2["d2h_bytes"], "seconds": t_e2e,
               "points_per_s": world * m / t_e2e, "note": "one plssvm_b200_predict call: H2D of the support vectors and the points from pinned memory + norms + tiles + D2H of the values"}
    if world > 1:
        tt = torch.tensor([wall, tile_ms], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall, tile_ms = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks_path = os.path.join(ROOT, "profiles", "peaks_b200.json")
    fp64_pipe_peak = float(json.load(open(peaks_path)).get("dmma_tflops_sustained_3s", 37.0)) if os.path.exists(peaks_path) else 37.0
    impl_used = t["impl_used"]
    peak, kname, peak_src = fp64_pipe_peak, "tile_kernel_dmma<rbf, rect>", "measured DMMA issue peak (profiles/peaks_b200.json)"
    if impl_used == 6:  # int8-slice tiles: int8 tensor pipe / 28 (see run_ours)
        i8_path = os.path.join(ROOT, "profiles", "r01", "i8_peaks_b200.json")
        i8 = json.load(open(i8_path)) if os.path.exists(i8_path) else {}
        k8 = "i8_mma_n256_random_tops_sustained_3s" if wall > 2.0 else "i8_mma_n256_random_tops_burst"
        peak = float(i8.get(k8, 4500.0)) / 28.0
        kname, peak_src = "tile_kernel_i8<rbf, rect> (fp64 through int8 slices, tcgen05 kind::i8)", f"int8 tensor pipe / 28: {k8} (tools/i8_peak_probe; profiles/r01/i8_peaks_b200.json)"
    achieved = F * args.steps / (tile_ms * 1e-3) / 1e12 if tile_ms > 0 else 0.0
    line = {
        "metric": "predict_tflops", "value": world * F * args.steps / wall / 1e12, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5: {desc}", "points_per_step_per_gpu": m, "support_vectors": n_sv, "features": d, "flops_per_step_per_gpu": F,
                   "l2": "points (2.1 GB) and support vectors (2.1 GB) larger than L2"},
        "points_per_s": world * m * args.steps / wall, "seconds_for_1048576_points": 1048576.0 / (world * m * args.steps / wall),
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None, "kernel": kname, "peak_source": peak_src,
                     "vs_fp64_pipe_peak": achieved / fp64_pipe_peak},
        "cpu_baseline": None, "tile_impl": int(impl_used),
    }
    print(json.dumps(_finite(line)), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _finite(obj):
    """JSON has no NaN / inf: iterations run far past convergence (fixed-count timing) may end with a non-finite residual."""
    if isinstance(obj, float):
        return obj if obj == obj and abs(obj) != float("inf") else None
    if isinstance(obj, dict):
        return {k: _finite(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_finite(v) for v in obj]
    return obj


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "C5":
        run_predict(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
