#!/usr/bin/env python
"""Benchmark of the hot path: the implicit kernel-matrix-vector product inside the CG solve of the reduced LS-SVM system.

    python bench.py --gpus N --steps K --warmup W            # our arm.  N > 1 under torchrun: one rank per GPU (the driver's launch);
                                                             # N > 1 WITHOUT torchrun: one process, one device group behind the C ABI
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own OpenMP kernels on the host cores

A *step* is one CG iteration (gpu_csvm.hpp:568-636): one implicit matvec Ad = Q~ d over the whole data set plus the vector
updates (every 50th iteration a second matvec recomputes the residual).  Metric = algorithmic matvec TFLOP/s inside the CG
loop: F = d * n * (n + 1) FLOPs per matvec x matvecs executed in the timed region / device time;
`cg_iters_per_s` = timed iterations / the same time.  Workload at every N: BASELINE.json configs[1],
65,536 x 4,096 dense, RBF gamma = 1/d, fp64 (strong scaling: tiles of the triangle are sharded over the ranks).

Timing: `value` — data resident in HBM, W untimed + exactly K timed iterations, device time from CUDA events recorded by
the library on its launching stream, bracketed by barrier + synchronize, max over ranks.  `e2e` — one
plssvm_b200_solve_f64 call on PINNED HOST buffers (upload of X and y, q-kernel, r0, K iterations, download of alpha).
`e2e_csvm` — the same through the reference's own `csvm::fit` with the b200 backend registered (integration/ref_bridge.cpp).
X (2.1 GB) is larger than L2 (126 MB) and streamed in full by every iteration, so no explicit L2 flush is needed.
With no flags the N = 1 run appends short runs of the other BASELINE configurations under "extra_workloads" (C1 full fit vs the CPU
reference, C3, C4 factorised, C5), each with its own roofline and clocks.  At N > 1 every line carries "parity_vs_n1": one sharded
matvec against the same matvec unsharded on rank 0 and against torch fp64 on 256 sampled rows, and whether all ranks hold the same alpha.
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
    ap.add_argument("--tile-impl", type=int, default=0, help="0 auto (int8 slices on tcgen05 kind::i8 for both real types: fp64 on single CTAs, fp32 on CTA pairs), 1 SIMT tiles, 2 fp64 DMMA / fp32 3xTF32 "
                    "tiles, 6 int8-slice tiles on single CTAs, 7 the same with 4 planes for fp32, 10 int8-slice tiles on CTA pairs (fp64: experimental builds only); "
                    "4 / 5 / 8 / 9 only in builds with -DPLSSVM_B200_EXPERIMENTAL")
    ap.add_argument("--no-dmma-line", action="store_true", help="fp64 only: skip the short extra run of the native-FP64 DMMA tiles reported under 'fp64_dmma_tiles'")
    ap.add_argument("--linear-factorized", action="store_true", help="linear kernel only: time the factorised X (X^T v) matvec (HBM-bound) instead of the implicit tiles")
    ap.add_argument("--full-solve", action="store_true", help="additionally run the whole fit to eps = 1e-8 (fp64) / 1e-4 (fp32) through the C ABI; with C1 also on the CPU reference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short extra workloads (C1 / C3 / C4 factorised / C5) the default N = 1 run appends")
    ap.add_argument("--no-through-csvm", action="store_true", help="skip the second end-to-end number through the reference's csvm::fit (integration/ref_bridge)")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=VALUE", help="extra plssvm_b200_set_option settings (development / A-B measurements)")
    ap.add_argument("--balance", type=int, default=1, help="several ranks: rate-weighted tile shares (1, default) or fixed equal shares (0)")
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
def _peaks():
    pk = {}
    for path in (os.path.join(ROOT, "profiles", "peaks_b200.json"), os.path.join(ROOT, "profiles", "r01", "i8_peaks_b200.json")):
        if os.path.exists(path):
            pk.update(json.load(open(path)))
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        pk["driver"] = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    return pk


def tile_roofline(impl_used, dtype, kernel, mode, achieved, avg_launch_ms, launches, flops_per_launch, traffic=None):
    """Roofline object of the dominant kernel (the tile kernel).  Denominators: tools/peak_probe + tools/i8_peak_probe on this pool's B200
    (profiles/peaks_b200.json, profiles/r01/i8_peaks_b200.json) — the pipe the kernel actually runs on; both the burst figure (a kernel timed
    alone at 1965 MHz) and the sustained one (seconds under the 1 kW cap) are reported on every line."""
    pk = _peaks()
    f64 = dtype == "float64"
    real = "fp64" if f64 else "fp32"
    fp_pipe = float(pk.get("dmma_tflops_sustained_3s", 37.0)) if f64 else float(pk.get("cublas_sgemm_tf32_random_tflops_sustained_4s", 770.0)) / 3.0
    out = {"bound": "tensor", "achieved": achieved, "unit": "TFLOP/s", "avg_launch_ms": avg_launch_ms, "launches_timed": int(launches), "flops_per_launch": flops_per_launch,
           "traffic": traffic}
    if impl_used in (6, 7, 8, 9, 10, 11):
        products = 28.0 if f64 else (10.0 if impl_used == 7 else 6.0)
        planes = 7 if f64 else (4 if impl_used == 7 else 3)
        sus, burst = float(pk.get("i8_mma_n256_random_tops_sustained_3s", 3819.0)) / products, float(pk.get("i8_mma_n256_random_tops_burst", 4425.0)) / products
        variant = {8: ", 2 x 2 CTA clusters + TMA multicast", 9: ", CTA pairs (cta_group::2)", 10: ", CTA pairs: cta_group::2, M = 256", 11: ", 2-CTA clusters sharing the A planes (TMA multicast)"}.get(impl_used, "")
        kname = "tile_kernel_i8_pair" if impl_used == 10 else "tile_kernel_i8"
        out.update(kernel=f"{kname}<{real}, {planes} int8 planes, {kernel}, {mode}> (tcgen05 kind::i8{variant})", peak=sus, frac=achieved / sus, peak_burst=burst,
                   frac_sustained=achieved / sus, frac_burst=achieved / burst, int8_tops=achieved * products, vs_float_pipe_peak=achieved / fp_pipe, float_pipe_peak_tflops=fp_pipe,
                   peak_source=f"int8 tensor pipe / {int(products)} int8 products per {real} product: tcgen05.mma kind::i8 issue-loop peak with random operands, sustained 3 s "
                               f"({pk.get('i8_mma_n256_random_tops_sustained_3s', 'nominal 3819')} TOPS) resp. burst ({pk.get('i8_mma_n256_random_tops_burst', 'nominal 4425')} TOPS), measured on "
                               "this pool's B200 by tools/i8_peak_probe (profiles/r01/i8_peaks_b200.json); `peak` / `frac` are the sustained ones")
    elif impl_used in (2, 4, 5):
        if f64:
            sus = burst = fp_pipe
            name, src = f"tile_kernel_dmma<{kernel}, {mode}> (TMA + mma.sync m8n8k4.f64)", "DMMA issue-loop peak measured by tools/peak_probe (profiles/peaks_b200.json; burst = sustained: the FP64 pipe is not power-limited)"
        else:
            sus, burst = fp_pipe, float(pk.get("cublas_sgemm_tf32_random_tflops_burst", 880.0)) / 3.0
            name = {2: "tile_kernel_tf32", 4: "tile_kernel_tf32_2sm (cta_group::2)", 5: "tile_kernel_tf32_n256"}[impl_used] + f"<{kernel}, {mode}> (tcgen05 3xTF32)"
            src = "cuBLAS TF32 GEMM 8192^3 with random operands / 3 (3xTF32 split), sustained 4 s resp. burst (profiles/peaks_b200.json)"
        out.update(kernel=name, peak=sus, frac=achieved / sus, peak_burst=burst, frac_sustained=achieved / sus, frac_burst=achieved / burst, peak_source=src)
    else:
        peak = float(pk.get("dfma_tflops", 37.0)) if f64 else float(pk.get("ffma_tflops", 70.0))
        out.update(kernel=f"tile_kernel_simt<{real}, {kernel}, {mode}>", bound="fma", peak=peak, frac=achieved / peak, peak_burst=peak, frac_sustained=achieved / peak,
                   frac_burst=achieved / peak, peak_source="DFMA / FFMA issue-loop peak (tools/peak_probe)")
    return out


class Ranks:
    """How this process takes part in the run: alone, as one rank of a torchrun launch (one process per GPU, torch.distributed over NCCL), or as the
    single process of an N-GPU run that drives a device group through the C ABI (plssvm_b200_create(device_ids, n_dev))."""

    def __init__(self, gpus):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.group = self.world == 1 and gpus > 1  # no torchrun: one process, device group
        self.n_gpus = gpus if self.group else self.world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=self.device)

    def backend(self, args):
        import plssvm_b200 as pb
        be = pb.Backend(devices=list(range(self.n_gpus))) if self.group else pb.Backend(self.local_rank)
        if args.tile_impl:
            be.set_option("impl", args.tile_impl)
        be.set_option("balance", args.balance)
        for kv in args.option:
            key, val = kv.split("=")
            be.set_option(key, int(val))
        if self.world > 1:
            be.init_comm_from_torch()
        return be

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        for i in range(self.n_gpus if self.group else 1):
            self.torch.cuda.synchronize(i if self.group else self.local_rank)

    def max(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def all_equal(self, array):
        """True iff every rank holds bit-identical values (device group: there is one result by construction)."""
        if self.world == 1:
            return True
        import numpy as np
        t = self.torch.from_numpy(np.ascontiguousarray(array).view(np.uint8).copy()).to(self.device)
        lo, hi = t.clone(), t.clone()
        self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN)
        self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX)
        return bool(self.torch.equal(lo, hi))

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def parity_vs_n1(rk, be, ds, X, kernel, dtype, args):
    """SCALE self-check (N > 1): ONE sharded matvec on the benchmark's own resident data against (a) the same matvec unsharded on this rank's GPU
    (a second, communicator-less context) and (b) torch fp64 (cuBLAS) on 256 sampled rows.  The run fails above 1e-12 (fp32: 2e-4)."""
    import numpy as np
    import plssvm_b200 as pb
    torch = rk.torch
    N, d = X.shape
    n = N - 1
    npdt = np.dtype(dtype)
    rng = np.random.default_rng(17)
    v = rng.uniform(1.0, 2.0, n).astype(npdt)
    q, k_last = be.run_q_kernel(ds, kernel)
    qa = float(k_last) + 1.0
    sharded = be.run_svm_kernel(ds, q, v, np.zeros(n, npdt), qa, 1.0, 1.0, kernel)
    rows = np.sort(rng.choice(n, 256, replace=False))
    rows_t = torch.from_numpy(rows).to(rk.device)
    A = X[rows_t].double()
    G = A @ X[:n].double().T
    if kernel == "linear":
        K = G
    elif kernel == "polynomial":
        K = ((1.0 / d) * G) ** 3
    else:
        sq = (X[:n].double() ** 2).sum(1)
        K = torch.exp(-(1.0 / d) * (sq[rows_t][:, None] + sq[None, :] - 2 * G).clamp_min(0))
    q_t, v_t = torch.from_numpy(q.astype(np.float64)).to(rk.device), torch.from_numpy(v.astype(np.float64)).to(rk.device)
    want = ((K + qa - q_t[rows_t][:, None] - q_t[None, :]) @ v_t + v_t[rows_t]).cpu().numpy()
    del A, G, K
    torch.cuda.empty_cache()
    err_torch = float(np.max(np.abs(sharded[rows].astype(np.float64) - want)) / np.max(np.abs(want)))
    err_n1 = None
    if rk.rank == 0:
        single = pb.Backend(rk.local_rank)
        if args.tile_impl:
            single.set_option("impl", args.tile_impl)
        ds1 = single.dataset(X)
        one = single.run_svm_kernel(ds1, q, v, np.zeros(n, npdt), qa, 1.0, 1.0, kernel)
        ds1.close()
        single.close()
        err_n1 = float(np.max(np.abs(sharded.astype(np.float64) - one.astype(np.float64))) / np.max(np.abs(one)))
    same = rk.all_equal(sharded)
    tol = 1e-12 if npdt == np.float64 else 2e-4
    out = {"max_rel_err": err_n1, "vs_torch_fp64_rows_max_rel_err": err_torch, "rows_sampled": 256, "matvec_equal_across_ranks": same, "tolerance": tol,
           "what": "one sharded implicit matvec on the benchmark's resident data vs the unsharded matvec on rank 0's GPU (all rows) and vs torch fp64 on 256 sampled rows"}
    ok = err_torch <= tol and same and (err_n1 is None or err_n1 <= tol)
    return out, ok


def through_csvm(args, N, d, kernel, dtype, Xh, yh, n_devices):
    """Second end-to-end number: the reference's OWN csvm::fit (data_set -> fit -> model) with the b200 backend registered, through
    integration/ref_bridge.cpp — the drop-in path a PLSSVM user takes (std::vector<std::vector<T>> rows, staged through the pinned ring)."""
    import ctypes
    import numpy as np
    path = os.path.join(ROOT, "integration", "_ref", "libplssvm_ref_bridge.so")
    if dtype != "float64" or not os.path.exists(path):
        return None
    try:
        import plssvm_b200
        plssvm_b200.load_library()
        lib = ctypes.CDLL(path)
        vp, sz, i32, f64, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double, ctypes.c_ulonglong
        lib.refb_last_error.restype = ctypes.c_char_p
        lib.refb_fit_timed_f64.argtypes = [i32, vp, sz, sz, vp, i32, i32, f64, f64, f64, f64, u64, vp, vp, vp]
        os.environ["PLSSVM_B200_NUM_DEVICES"] = str(n_devices)
        X = Xh.numpy()
        labels = np.where(yh.numpy() > 0, 1, -1).astype(np.int32)
        alpha, rho, secs = np.empty(N), np.zeros(1), np.zeros(2)
        rc = lib.refb_fit_timed_f64(1, X.ctypes.data, N, d, labels.ctypes.data, KERNEL_IDS[kernel], 3, 0.0, 0.0, 1.0, 1e-30, args.steps, alpha.ctypes.data, rho.ctypes.data,
                                    secs.ctypes.data)
        if rc != 0:
            return {"error": lib.refb_last_error().decode(errors="replace")[:300]}
        matvecs = args.steps + 1 + args.steps // 50
        return {"value": matvec_flops(N, d) * matvecs / secs[0] / 1e12, "unit": "TFLOP/s", "seconds": float(secs[0]), "data_set_seconds": float(secs[1]), "matvecs": matvecs,
                "iterations": args.steps, "devices": n_devices, "h2d_bytes_per_step": (N * d + N) * 8 / args.steps, "d2h_bytes_per_step": N * 8 / args.steps,
                "note": "plssvm::csvm::fit of the reference (unmodified base class) on plssvm::b200x::csvm: the backend gets std::vector<std::vector<double>> rows and stages them "
                        "through its pinned ring; `seconds` is the fit() call alone, `data_set_seconds` the reference's own construction of the data_set before it"}
    except Exception as e:  # optional leg: never fail the benchmark because of it
        return {"error": str(e)[:300]}


def measure_cg(rk, be, args, workload, steps, warmup, N=None, d=None, factorized=False, want_e2e=False, want_parity=False):
    """W warm-up + exactly K CG iterations on resident data; returns the measurement dict (value, ms_per_step, roofline, clocks, ...)."""
    import numpy as np
    torch = rk.torch
    N0, d0, kernel, dtype, desc = WORKLOADS[workload]
    N, d = N or N0, d or d0
    F = matvec_flops(N, d)
    npdt = np.dtype(dtype)
    X, y = make_device_data(N, d, dtype, 42 + list(WORKLOADS).index(workload), rk.device)
    y_host = y.cpu().numpy()
    ds = be.dataset(X)
    out = {"workload": workload, "kernel": kernel, "dtype": dtype, "N": N, "d": d, "desc": desc, "F": F}
    if want_parity and rk.n_gpus > 1:
        out["parity_vs_n1"], out["parity_ok"] = parity_vs_n1(rk, be, ds, X, kernel, dtype, args)
    eps = 1e-30 if dtype == "float64" else 1e-18  # never met: the iteration count is fixed
    be.set_option("ignore_convergence", 1)  # fp32 CG can hit an exactly-zero residual after ~15 iterations on this data; time exactly K iterations
    be.set_option("linear_factorized", 1 if factorized else 0)
    cg = be.cg_begin(ds, y_host, kernel, eps=eps)
    cg.step(warmup)
    t_before = be.timings()
    rk.barrier()
    with ClockSampler(rk.local_rank) as clocks:
        wall0 = time.perf_counter()
        done_iters, _ = cg.step(steps)
        rk.barrier()
        wall = time.perf_counter() - wall0
    t_after = be.timings()
    res = cg.finish()
    be.set_option("ignore_convergence", 0)
    be.set_option("linear_factorized", 0)
    assert done_iters == warmup + steps, (done_iters, warmup, steps)
    dev_ms = t_after["cg_loop_ms"] - t_before["cg_loop_ms"]
    tile_ms = t_after["matvec_tile_ms"] - t_before["matvec_tile_ms"]
    mv_ms = t_after["matvec_ms"] - t_before["matvec_ms"]
    calls = t_after["matvec_calls"] - t_before["matvec_calls"]
    dev_ms, wall = rk.max(dev_ms, wall)
    out.update(dev_ms=dev_ms, wall=wall, tile_ms=tile_ms, matvec_ms=mv_ms, calls=int(calls), launches=int(t_after["kernel_launches"] - t_before["kernel_launches"]),
               impl_used=int(t_after["impl_used"]), clocks=clocks.summary(), final_residual=float(res["delta"]), rebalances=int(t_after["rebalances"]),
               value=F * calls / (dev_ms * 1e-3) / 1e12, ms_per_step=dev_ms / steps, alpha_equal_across_ranks=rk.all_equal(res["alpha"]))
    if want_e2e:
        # end to end through the C ABI with pinned host buffers: one plssvm_b200_solve call (upload, q, r0, K iterations, download)
        Xh = torch.empty((N, d), dtype=X.dtype, pin_memory=True)
        Xh.copy_(X)
        yh = torch.empty(N, dtype=X.dtype, pin_memory=True)
        yh.copy_(y)
        ds.close()
        del X
        torch.cuda.empty_cache()
        be.set_option("ignore_convergence", 1)
        rk.barrier()
        t0 = time.perf_counter()
        r2 = be.solve(Xh, yh, kernel, eps=eps, max_iter=steps)
        rk.barrier()
        (t_e2e,) = rk.max(time.perf_counter() - t0)
        be.set_option("ignore_convergence", 0)
        t2 = be.timings()
        out["e2e"] = {"value": F * t2["matvec_calls"] / t_e2e / 1e12, "unit": "TFLOP/s", "matvecs": int(t2["matvec_calls"]), "cg_iters_per_s": r2["iterations"] / t_e2e,
                      "h2d_bytes_per_step": t2["h2d_bytes"] / r2["iterations"], "d2h_bytes_per_step": t2["d2h_bytes"] / r2["iterations"], "seconds": t_e2e,
                      "iterations": r2["iterations"],
                      "note": "one plssvm_b200_solve call: H2D of X and y from pinned memory (several GPUs: every GPU uploads 1 / N of the rows, NCCL all-gather) + q-kernel + r0 "
                              "matvec + K iterations + D2H of alpha; bytes are summed over the GPUs, per call / K"}
        out["host"] = (Xh, yh)
    else:
        ds.close()
        del X
        torch.cuda.empty_cache()
    return out


def cg_roofline(m, n_gpus):
    """Roofline of the dominant kernel of a measure_cg() result."""
    import numpy as np
    avg_tile_s = m["tile_ms"] / max(m["calls"], 1) * 1e-3
    if m["impl_used"] == 3:  # factorised linear matvec: HBM-bound, two passes over X per matvec, whole matvec against the measured copy bandwidth
        pk = _peaks()
        hbm = float(pk.get("driver", {}).get("hbm_gbs", 6650.0))
        mv_s = m["matvec_ms"] / max(m["calls"], 1) * 1e-3
        gbs = 2.0 * (m["N"] - 1) * m["d"] * np.dtype(m["dtype"]).itemsize / mv_s / 1e9
        return {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None, "kernel": "w_partial_kernel + linear_fact_apply_kernel (X read twice)",
                "avg_matvec_ms": mv_s * 1e3, "peak_source": "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)" if "driver" in pk else "fallback 6650 GB/s"}
    achieved = (m["F"] / n_gpus) / avg_tile_s / 1e12 if avg_tile_s > 0 else 0.0
    traffic = None
    ncu_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(ncu_path) and m["N"] == WORKLOADS[m["workload"]][0] and n_gpus == 1:
        traffic = json.load(open(ncu_path)).get(m["workload"] + ("_i8" if m["impl_used"] in (6, 7, 10) else ""))
    return tile_roofline(m["impl_used"], m["dtype"], m["kernel"], "sym", achieved, avg_tile_s * 1e3, m["calls"], m["F"] / n_gpus, traffic)


def extra_workloads(rk, be, args):
    """Short runs of the other BASELINE configurations so that every config is measured by the default driver run (N = 1 only)."""
    import numpy as np
    extras = []

    def guarded(name, fn):
        try:
            t0 = time.perf_counter()
            e = fn()
            e["seconds_spent"] = time.perf_counter() - t0
            extras.append(e)
        except Exception as ex:  # an extra must never take the headline line down
            extras.append({"workload": name, "error": str(ex)[:300]})

    def cg_entry(workload, steps, warmup, factorized=False):
        m = measure_cg(rk, be, args, workload, steps, warmup, factorized=factorized)
        N, d = m["N"], m["d"]
        return {"workload": f"{workload}: {m['desc']}" + (" [factorised linear fast path, SURVEY 8f row 4]" if factorized else ""), "metric": "cg_matvec_tflops", "value": m["value"],
                "unit": "TFLOP/s", "steps": steps, "warmup": warmup, "ms_per_step": m["ms_per_step"], "cg_iters_per_s": steps / (m["dev_ms"] * 1e-3), "dtype": "f64" if m["dtype"] == "float64" else "f32",
                "tile_impl": m["impl_used"], "gpu_launches": m["launches"], "roofline": cg_roofline(m, 1), "clocks": m["clocks"], "rows": N, "features": d}

    def c1_full_fit():
        import oracle
        N, d, kernel, dtype, desc = WORKLOADS["C1"]
        X, y = make_host_data(N, d, dtype, 42)
        t0 = time.perf_counter()
        r = be.solve(X, y, kernel, eps=1e-8)
        gpu_s = time.perf_counter() - t0  # the FIRST linear-kernel call of this context: includes loading the kernels and growing the workspaces
        t0 = time.perf_counter()
        be.solve(X, y, kernel, eps=1e-8)
        gpu_repeat_s = time.perf_counter() - t0
        t = be.timings()
        orc = oracle.Oracle("reference" if oracle.available("reference") else "port")
        orc.set_threads(len(os.sched_getaffinity(0)))
        t0 = time.perf_counter()
        rc = orc.solve(KERNEL_IDS[kernel], X, y, gamma=1.0 / d, eps=1e-8)
        cpu_s = time.perf_counter() - t0
        same = r["iterations"] == rc["iterations"]
        return {"workload": f"C1: {desc}, whole fit to eps = 1e-8 through plssvm_b200_solve_f64 (host buffers) vs the reference's OpenMP path", "metric": "fit_seconds", "value": gpu_s,
                "unit": "s", "higher_is_better": False, "repeat_call_seconds": gpu_repeat_s, "cg_loop_ms": t["cg_loop_ms"], "iterations": r["iterations"], "cpu_seconds": cpu_s, "cpu_iterations": rc["iterations"], "cpu_threads": orc.max_threads(),
                "cpu_kind": orc.reported_kind(), "speedup_vs_cpu": cpu_s / gpu_s, "tile_impl": int(t["impl_used"]), "gpu_launches": int(t["kernel_launches"]),
                "alpha_max_rel_diff_vs_cpu": float(np.max(np.abs(r["alpha"] - rc["alpha"])) / np.max(np.abs(rc["alpha"]))) if same else None,
                "matvec_tflops_incl_setup": matvec_flops(N, d) * t["matvec_calls"] / (t["matvec_ms"] * 1e-3) / 1e12 if t["matvec_ms"] > 0 else None}

    guarded("C1", c1_full_fit)
    guarded("C3", lambda: cg_entry("C3", 5, 3))
    guarded("C4_factorized", lambda: cg_entry("C4", 5, 3, factorized=True))
    guarded("C5", lambda: predict_measure(rk, be, args, steps=1, warmup=1, want_e2e=False))
    return extras


def run_ours(args):
    import numpy as np

    rk = Ranks(args.gpus)
    be = rk.backend(args)
    N0, d0, kernel, dtype, desc = WORKLOADS[args.workload]
    headline = not (args.rows or args.features)
    m = measure_cg(rk, be, args, args.workload, args.steps, args.warmup, N=args.rows or None, d=args.features or None, factorized=args.linear_factorized,
                   want_e2e=not args.no_e2e, want_parity=True)
    N, d, F = m["N"], m["d"], m["F"]
    e2e = m.get("e2e")

    # ---- optional: the whole fit to the parity tolerance, GPU vs the reference's CPU path (C1 is run to convergence)
    full = None
    if args.full_solve and not args.no_e2e:
        Xh, yh = m["host"]
        feps = 1e-8 if dtype == "float64" else 1e-4
        rk.barrier()
        t0 = time.perf_counter()
        rf = be.solve(Xh, yh, kernel, eps=feps)
        rk.barrier()
        full = {"eps": feps, "gpu_seconds": time.perf_counter() - t0, "gpu_iterations": rf["iterations"]}

    # ---- second end-to-end number: the reference's own csvm::fit on the b200 backend (one process, all N devices) — not under torchrun
    e2e_csvm = None
    if rk.world == 1 and e2e is not None and not args.no_through_csvm and headline is not None:
        Xh, yh = m["host"]
        e2e_csvm = through_csvm(args, N, d, kernel, dtype, Xh, yh, rk.n_gpus)
    m.pop("host", None)

    if rk.rank != 0:
        rk.close()
        return

    roofline = cg_roofline(m, rk.n_gpus)
    impl_used = m["impl_used"]

    # fp64: the native-FP64 DMMA tiles (north-star kernel, tile_dmma.cuh) on the same workload, a few iterations, against the DMMA peak
    dmma_line = None
    if dtype == "float64" and impl_used == 6 and rk.n_gpus == 1 and not args.no_dmma_line and not args.linear_factorized:
        be.set_option("impl", 2)
        m2 = measure_cg(rk, be, args, args.workload, 3, 1, N=N, d=d)
        be.set_option("impl", args.tile_impl)
        r2 = cg_roofline(m2, 1)
        dmma_line = {"kernel": r2["kernel"], "avg_launch_ms": r2["avg_launch_ms"], "achieved": r2["achieved"], "peak": r2["peak"], "unit": "TFLOP/s", "frac": r2["frac"],
                     "traffic": r2["traffic"], "launches_timed": r2["launches_timed"],
                     "note": "the same matvec with --tile-impl 2: TMA + mma.sync m8n8k4.f64 tiles against the measured DMMA issue peak (the kernel the north star names)"}

    cpu_baseline = None
    if rk.n_gpus == 1 and not args.no_cpu_baseline:
        orc, kind, cores, n_rows, run = cpu_matvec_sample("reference", d, kernel, dtype, args.cpu_seconds)
        t_cpu = run(n_rows)
        cpu_baseline = {"value": d * float(n_rows) * (n_rows + 1) / t_cpu / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "reference" if kind == "reference" else "port",
                        "sample": f"one OpenMP matvec on the leading {n_rows} rows x {d} features ({t_cpu:.1f} s); same kernel / real type"}

    # optional second baseline: the reference's own CUDA kernel on this GPU, on a slice of the workload
    ref_cuda = None
    if rk.n_gpus == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            if oracle.RefCuda.available():
                npdt = np.dtype(dtype)
                rows = min(N, 8192 + 1)
                Xs, _ = make_device_data(rows, d, dtype, 4242, rk.device)
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

    extras = None
    if rk.n_gpus == 1 and headline and args.workload == "C2" and not args.no_extra and not args.linear_factorized:
        extras = extra_workloads(rk, be, args)

    planes = {6: 7 if dtype == "float64" else 3, 7: 7 if dtype == "float64" else 4, 10: 7 if dtype == "float64" else 3}.get(impl_used)
    line = {
        "metric": "cg_matvec_tflops", "value": m["value"], "unit": "TFLOP/s", "n_gpus": rk.n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if dtype == "float64" else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}" + ("" if headline else f" [DEV OVERRIDE rows={N} features={d}]") + (" [factorised linear fast path]" if args.linear_factorized else ""),
                   "kernel": kernel, "rows": N, "features": d, "flops_per_step": F, "l2": f"inputs ({N * d * (8 if dtype == 'float64' else 4) / 1e9:.1f} GB of X, {N * d * (7 if dtype == 'float64' else 3) / 1e9:.1f} GB of digit planes) larger than the 126 MB L2; no flush needed",
                   "parallelism": f"triangle tiles sharded over {rk.n_gpus} GPU(s) ({'one process, device group behind the C ABI' if rk.group else 'one process per GPU' if rk.world > 1 else 'single GPU'}), "
                                  f"X replicated, {'rate-weighted' if args.balance and rk.n_gpus > 1 else 'equal'} tile shares"},
        "cg_iters_per_s": args.steps / (m["dev_ms"] * 1e-3), "wall_ms_per_step": m["wall"] / args.steps * 1e3,
        "clocks": m["clocks"], "e2e": e2e, "e2e_csvm": e2e_csvm, "gpu_launches": m["launches"], "roofline": roofline, "cpu_baseline": cpu_baseline,
        "parity_vs_n1": (dict(m["parity_vs_n1"], alpha_equal_across_ranks=m["alpha_equal_across_ranks"]) if "parity_vs_n1" in m else None),
        "final_residual": m["final_residual"], "tile_impl": impl_used, "tile_share_rebalances": m["rebalances"], "full_solve": full,
        "precision_note": (f"{'fp64' if dtype == 'float64' else 'fp32'} storage, vectors and epilogue; x_i.x_j through {planes} int8 digit planes per operand, exact int32 tensor-core "
                           "products recombined in fp64 (rbf on data centred at the feature means); error vs the extended-precision oracle at or below the reference's own "
                           "(profiles/r02/parity_report_*.json)" if planes else
                           ("native fp64 DMMA tiles" if dtype == "float64" and impl_used == 2 else "fp32 storage and accumulation; products via the 3xTF32 split on tcgen05" if impl_used == 2 else None)),
        "matvecs_in_timed_region": m["calls"], "reference_cuda_baseline": ref_cuda, "fp64_dmma_tiles": dmma_line, "extra_workloads": extras,
    }
    print(json.dumps(_finite(line)), flush=True)
    ok = m.get("parity_ok", True) and m["alpha_equal_across_ranks"]
    rk.close()
    if not ok:
        raise SystemExit("bench.py: the sharded run failed its parity check against the unsharded matvec (see parity_vs_n1)")


# ---- prediction workload (C5): run_predict_kernel throughput -----------------------------------------------------------------------
def predict_measure(rk, be, args, steps, warmup, want_e2e):
    """A step = decision values of PREDICT_STEP_POINTS test points PER GPU against all support vectors (2 m n_sv d FLOPs per GPU); with several
    GPUs the library shards the n_gpus x m points of a step by ranges (no data-path collective)."""
    import numpy as np
    torch = rk.torch
    n_sv, d, kernel, dtype, desc = WORKLOADS["C5"]
    n_sv, d = args.rows or n_sv, args.features or d
    m = PREDICT_STEP_POINTS
    G = rk.n_gpus
    F = 2.0 * m * n_sv * d  # per GPU and step
    SV, _ = make_device_data(n_sv, d, dtype, 47, rk.device)
    P, _ = make_device_data(G * m, d, dtype, 48, rk.device)
    rng = np.random.default_rng(47)
    alpha = rng.uniform(-1, 1, n_sv)
    alpha -= alpha.mean()
    rho = 0.1
    sv_ds, p_ds = be.dataset(SV), be.dataset(P)
    for _ in range(warmup):
        be.predict_values(sv_ds, alpha, rho, p_ds, kernel)
    rk.barrier()
    tile_ms, launches = 0.0, 0
    with ClockSampler(rk.local_rank) as clocks:
        t0 = time.perf_counter()
        for _ in range(steps):
            vals, _ = be.predict_values(sv_ds, alpha, rho, p_ds, kernel)
            t = be.timings()
            tile_ms += t["matvec_tile_ms"]
            launches += t["kernel_launches"]
        rk.barrier()
        wall = time.perf_counter() - t0
    e2e = None
    if want_e2e:
        Ph = torch.empty((G * m, d), dtype=P.dtype, pin_memory=True)
        Ph.copy_(P)
        SVh = torch.empty((n_sv, d), dtype=P.dtype, pin_memory=True)
        SVh.copy_(SV)
        rk.barrier()
        t0 = time.perf_counter()
        vals2, _ = be.predict_values(SVh, alpha, rho, Ph, kernel)
        rk.barrier()
        (t_e2e,) = rk.max(time.perf_counter() - t0)
        t2 = be.timings()
        assert np.array_equal(vals, vals2)
        e2e = {"value": G * F / t_e2e / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": t2["h2d_bytes"], "d2h_bytes_per_step": t2["d2h_bytes"], "seconds": t_e2e,
               "points_per_s": G * m / t_e2e, "note": "one plssvm_b200_predict call: H2D of the support vectors (1 / N per GPU + all-gather) and of each GPU's range of points from pinned "
                                                      "memory + norms + digit planes + tiles + D2H of the values"}
    sv_ds.close()
    p_ds.close()
    del SV, P
    torch.cuda.empty_cache()
    wall, tile_ms = rk.max(wall, tile_ms)
    impl_used = int(t["impl_used"])
    achieved = F * steps / (tile_ms * 1e-3) / 1e12 if tile_ms > 0 else 0.0
    return {
        "metric": "predict_tflops", "value": G * F * steps / wall / 1e12, "unit": "TFLOP/s", "n_gpus": G, "steps": steps, "warmup": warmup,
        "ms_per_step": wall / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "workload": f"C5: {desc}",
        "config": {"workload": f"C5: {desc}", "points_per_step_per_gpu": m, "support_vectors": n_sv, "features": d, "flops_per_step_per_gpu": F,
                   "l2": "points (2.1 GB per GPU) and support vectors (2.1 GB) larger than L2", "parallelism": f"test points sharded by ranges over {G} GPU(s) inside the library, no collective"},
        "points_per_s": G * m * steps / wall, "seconds_for_1048576_points": 1048576.0 / (G * m * steps / wall),
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
        "roofline": tile_roofline(impl_used, dtype, kernel, "rect", achieved, tile_ms / max(steps, 1), steps, F),
        "cpu_baseline": None, "tile_impl": impl_used,
    }


def run_predict(args):
    rk = Ranks(args.gpus)
    be = rk.backend(args)
    line = predict_measure(rk, be, args, args.steps, args.warmup, not args.no_e2e)
    if rk.rank == 0:
        print(json.dumps(_finite(line)), flush=True)
    rk.close()


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
