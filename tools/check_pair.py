#!/usr/bin/env python
"""Hardware check of the CTA-pair wide-N int8-slice kernel (impl 10, tile_i8_pair.cuh) against the single-CTA kernel (impl 6):
bit-identity of matvec / predict results over ragged shapes for both real types, then tile-kernel time A/B on the bench shapes.
    python tools/check_pair.py [--quick] [--no-time] [--skip-parity] [--only=C2|C3] [--kernel=linear|polynomial|rbf] [--stats]
(fp64 runs the pair kernel only in a library built with PLSSVM_B200_EXPERIMENTAL=1, otherwise impl 10 resolves to 6; PLSSVM_B200_LIB selects the library:
A/B of two builds on one box.  --stats prints the per-role wait fractions; with a -DPB_TILE_STATS_FINE build also the split of the epilogue's time.)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import plssvm_b200 as pb  # noqa: E402
from bench import WORKLOADS, make_device_data, matvec_flops  # noqa: E402
from datagen import make_data  # noqa: E402

A, B = 6, int(([a.split("=")[1] for a in sys.argv if a.startswith("--b=")] or ["10"])[0])  # --b=11: the A-sharing 2-CTA cluster variant (experimental builds)
be = pb.Backend(0)
dev = torch.device("cuda", 0)
quick = "--quick" in sys.argv
only = [a.split("=")[1] for a in sys.argv if a.startswith("--only=")]  # --only=C2 / --only=C3: one real type
dtypes = {"C2": (np.float64,), "C3": (np.float32,)}[only[0]] if only else (np.float64, np.float32)
bad = 0
for dtype in (() if "--skip-parity" in sys.argv else dtypes):
    for (N, d) in ((2, 3), (130, 1), (257, 40), (386, 65), (1000, 96), (2049, 333), (700, 1200), (5000, 129)):
        for kernel in ("linear", "polynomial", "rbf"):
            X, y = make_data(N, d, 7, dtype)
            if N > 300:
                X[3] *= 1e-3
                X[5] *= 30.0
            n = N - 1
            ds = be.dataset(X)
            q, k_last = be.run_q_kernel(ds, kernel)
            v = np.random.default_rng(3).uniform(1, 2, n).astype(dtype)
            outs = {}
            for impl in (A, B):
                be.set_option("impl", impl)
                outs[impl] = be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
                assert be.timings()["impl_used"] == impl, be.timings()["impl_used"]
            be.set_option("impl", 0)
            same = bool(np.array_equal(outs[A], outs[B]))
            bad += 0 if same else 1
            rel = float(np.max(np.abs(outs[A] - outs[B])) / np.max(np.abs(outs[A])))
            print(f"{np.dtype(dtype).name} N={N:5d} d={d:4d} {kernel:10s} matvec identical {same}  rel diff {rel:.2e}", flush=True)
            ds.close()
    X, y = make_data(3000, 200, 11, dtype)
    P, _ = make_data(700, 200, 10, dtype)
    alpha = np.random.default_rng(5).standard_normal(3000).astype(dtype)
    for kernel in ("polynomial", "rbf"):
        vals = {}
        for impl in (A, B):
            be.set_option("impl", impl)
            vals[impl], _ = be.predict_values(be.dataset(X), alpha, 0.1, be.dataset(P), kernel)
            vals[(impl, "host")], _ = be.predict_values(X, alpha, 0.1, P, kernel)
        be.set_option("impl", 0)
        same = bool(np.array_equal(vals[A], vals[B])) and bool(np.array_equal(vals[(A, "host")], vals[(B, "host")]))
        bad += 0 if same else 1
        print(f"{np.dtype(dtype).name} predict {kernel}: identical {same}", flush=True)
    # a whole solve
    X, y = make_data(1500, 128, 21, dtype)
    res = {}
    for impl in (A, B):
        be.set_option("impl", impl)
        res[impl] = be.solve(X, y, "rbf", eps=1e-6)
    be.set_option("impl", 0)
    same = bool(np.array_equal(res[A]["alpha"], res[B]["alpha"])) and res[A]["iterations"] == res[B]["iterations"]
    bad += 0 if same else 1
    print(f"{np.dtype(dtype).name} solve rbf 1500x128: iterations {res[A]['iterations']} / {res[B]['iterations']}  alpha identical {same}", flush=True)
print("PARITY", "OK" if bad == 0 else f"FAIL ({bad} cases)", flush=True)
if "--skip-parity" in sys.argv:
    pass
elif bad != 0 or "--no-time" in sys.argv:
    sys.exit(1 if bad else 0)

shapes = [("C2", 16384, 4096), ("C3", 32768, 1024)] if quick else [("C2", 16384, 4096), ("C2", 65536, 4096), ("C3", 32768, 1024), ("C3", 131072, 1024)]
for workload, rows, feats in shapes:
    if only and workload != only[0]:
        continue
    _, _, kernel, dtype, _ = WORKLOADS[workload]
    kernel = ([a.split("=")[1] for a in sys.argv if a.startswith("--kernel=")] or [kernel])[0]  # --kernel=linear: the same shape with another kernel function
    X, _ = make_device_data(rows, feats, dtype, 7, dev)
    ds = be.dataset(X)
    q, k_last = be.run_q_kernel(ds, kernel)
    v = np.random.default_rng(1).uniform(1, 2, rows - 1).astype(np.dtype(dtype))
    out = {"workload": workload, "rows": rows, "features": feats, "kernel": kernel}
    res = {}
    reps = 3 if rows <= 32768 else 6
    for impl in (A, B, A, B):
        be.set_option("impl", impl)
        ms = []
        for _ in range(reps):
            res[impl] = be.run_svm_kernel(ds, q, v, np.zeros_like(v), float(k_last) + 1.0, 1.0, 1.0, kernel)
            ms.append(be.timings()["matvec_tile_ms"])
        out.setdefault(f"tflops_impl{impl}", []).append([round(matvec_flops(rows, feats) / (m * 1e-3) / 1e12, 2) for m in ms])
    be.set_option("impl", 0)
    out["bit_identical"] = bool(np.array_equal(res[A], res[B]))
    print(json.dumps(out), flush=True)
    if "--stats" in sys.argv:
        be.set_option("tile_stats", 2)
        for impl in (A, B):
            be.set_option("impl", impl)
            be.run_svm_kernel(ds, q, v, np.zeros_like(v), float(k_last) + 1.0, 1.0, 1.0, kernel)
            t = be.timings()
            print(json.dumps({"impl": impl, "workload": workload, "rows": rows, **{k: round(float(t[k]), 4) for k in t if k.startswith("tile_")}}), flush=True)
        be.set_option("tile_stats", 0)
        be.set_option("impl", 0)
    ds.close()
    del X
    torch.cuda.empty_cache()
