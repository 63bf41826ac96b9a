"""Hardware check of the fp64 int8-slice tensor path (tcgen05 kind::i8, tile_i8.cuh) against the DMMA tiles and the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import plssvm_b200 as pb  # noqa: E402
from datagen import make_data  # noqa: E402

be = pb.Backend(0)
orc = oracle.Oracle("reference" if oracle.available("reference") else "port")
quick = "--quick" in sys.argv
worst = 0.0
for (N, d) in ((2, 3), (130, 1), (130, 40), (300, 64), (386, 65), (1000, 96), (2049, 333), (700, 1200)):
    for kernel, kid in (("linear", 0), ("polynomial", 1), ("rbf", 2)):
        X, y = make_data(N, d, 7, np.float64)
        if N > 100:  # rows of very different magnitude exercise the per-row scaling
            X[3] *= 1e-6
            X[5] *= 300.0
            X[7, ::2] *= 1e-9
        n = N - 1
        ds = be.dataset(X)
        q, k_last = be.run_q_kernel(ds, kernel)
        v = np.random.default_rng(3).uniform(1, 2, n)
        outs = {}
        for impl in (2, 6, 8):
            be.set_option("impl", impl)
            outs[impl] = be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
        be.set_option("impl", 0)
        want = orc.matvec(kid, X, q, v, np.zeros(n), float(k_last) + 1.0, 1.0, 1.0, gamma=1.0 / d)
        sc = np.max(np.abs(want))
        e2 = np.max(np.abs(outs[2] - want)) / sc
        e6 = np.max(np.abs(outs[6] - want)) / sc
        e8 = np.max(np.abs(outs[8] - want)) / sc
        worst = max(worst, e6, e8)
        print(f"N={N:5d} d={d:4d} {kernel:10s} dmma {e2:.2e}  i8 {e6:.2e}  i8-cluster {e8:.2e}  bit-identical {np.array_equal(outs[6], outs[8])}  (vs fp64 oracle)", flush=True)
print("worst i8 matvec error:", worst, "OK" if worst < 1e-12 else "FAIL", flush=True)

# predict (rectangular tiles)
X, y = make_data(3000, 200, 11, np.float64)
ds = be.dataset(X)
P, _ = make_data(700, 200, 10, np.float64)
alpha = np.random.default_rng(5).standard_normal(3000)
for kernel in ("polynomial", "rbf"):
    vals = {}
    for impl in (2, 6, 8):
        be.set_option("impl", impl)
        vals[impl], _ = be.predict_values(ds, alpha, 0.1, be.dataset(P), kernel)
        vals[(impl, "host")], _ = be.predict_values(X, alpha, 0.1, P, kernel)
    sc = np.max(np.abs(vals[2]))
    print(f"predict {kernel}: |i8 - dmma| / scale = {np.max(np.abs(vals[6] - vals[2])) / sc:.2e}   host-staged: {np.max(np.abs(vals[(6, 'host')] - vals[2])) / sc:.2e}"
          f"   cluster: {np.max(np.abs(vals[8] - vals[2])) / sc:.2e} / {np.max(np.abs(vals[(8, 'host')] - vals[2])) / sc:.2e}", flush=True)
be.set_option("impl", 0)

# full solve parity between the two tile kernels
X, y = make_data(1500, 128, 21, np.float64)
res = {}
for impl in (2, 6):
    be.set_option("impl", impl)
    res[impl] = be.solve(X, y, "rbf", eps=1e-10)
a2, a6 = res[2]["alpha"], res[6]["alpha"]
print("solve rbf 1500x128: iterations", res[2]["iterations"], res[6]["iterations"], " alpha rel diff", float(np.max(np.abs(a2 - a6)) / np.max(np.abs(a2))), " rho", res[2]["rho"],
      res[6]["rho"], flush=True)
be.set_option("impl", 0)

# timing
for (N, d, kernel) in ((16385, 4096, "rbf"),) if quick else ((16385, 4096, "rbf"), (32769, 4096, "rbf"), (32769, 1024, "polynomial"), (32769, 2048, "linear")):
    X, y = make_data(N, d, 9, np.float64)
    ds = be.dataset(X)
    q, k_last = be.run_q_kernel(ds, kernel)
    v = np.ones(N - 1)
    for impl in (2, 6, 8):
        be.set_option("impl", impl)
        ts = []
        for _ in range(4):
            out = be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
            t = be.timings()
            ts.append(t["matvec_tile_ms"])
        print(f"{N}x{d} {kernel} impl {impl}: tile kernel {min(ts):.3f} ms -> {t['matvec_flops'] / min(ts) / 1e9:.1f} TFLOP/s (fp64-equivalent)  all: {[round(x, 3) for x in ts]}", flush=True)
        if impl == 2:
            ref = out
        else:
            print("   |i8 - dmma| / scale =", float(np.max(np.abs(out - ref)) / np.max(np.abs(ref))), flush=True)
    be.set_option("impl", 0)
    del ds

# ---- fp32 through int8 slices (S = 4) vs the 3xTF32 tiles -------------------------------------------------------------------------
worst32 = 0.0
for (N, d) in ((2, 3), (130, 1), (386, 65), (1000, 96), (2049, 333)):
    for kernel, kid in (("linear", 0), ("polynomial", 1), ("rbf", 2)):
        X, y = make_data(N, d, 7, np.float32)
        if N > 300:
            X[3] *= 1e-3
            X[5] *= 30.0
        n = N - 1
        ds = be.dataset(X)
        q, k_last = be.run_q_kernel(ds, kernel)
        v = np.random.default_rng(3).uniform(1, 2, n).astype(np.float32)
        outs = {}
        for impl in (1, 2, 6, 8, 9):
            be.set_option("impl", impl)
            outs[impl] = be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
        be.set_option("impl", 0)
        want = orc.matvec(kid, X.astype(np.float64), q.astype(np.float64), v.astype(np.float64), np.zeros(n), float(k_last) + 1.0, 1.0, 1.0, gamma=1.0 / d)
        sc = np.max(np.abs(want))
        e = {i: np.max(np.abs(outs[i] - want)) / sc for i in outs}
        worst32 = max(worst32, e[6], e[8], e[9])
        print(f"fp32 N={N:5d} d={d:4d} {kernel:10s} simt {e[1]:.2e}  tf32x3 {e[2]:.2e}  i8 {e[6]:.2e}  i8-cluster {e[8]:.2e}  i8-pair {e[9]:.2e} identical {np.array_equal(outs[6], outs[9])}  (vs fp64 oracle)", flush=True)
print("worst fp32 i8 matvec error:", worst32, "OK" if worst32 < 2e-6 else "FAIL", flush=True)
X, y = make_data(3000, 200, 11, np.float32)
P, _ = make_data(700, 200, 10, np.float32)
alpha = np.random.default_rng(5).standard_normal(3000).astype(np.float32)
for kernel in ("polynomial", "rbf"):
    vals = {}
    for impl in (2, 6, 9):
        be.set_option("impl", impl)
        vals[impl], _ = be.predict_values(be.dataset(X), alpha, 0.1, be.dataset(P), kernel)
        vals[(impl, "host")], _ = be.predict_values(X, alpha, 0.1, P, kernel)
    sc = np.max(np.abs(vals[2]))
    print(f"fp32 predict {kernel}: |i8 - tf32x3| / scale = {np.max(np.abs(vals[6] - vals[2])) / sc:.2e}   host-staged: {np.max(np.abs(vals[(6, 'host')] - vals[2])) / sc:.2e}"
          f"   pair identical: {np.array_equal(vals[6], vals[9])} / {np.array_equal(vals[(6, 'host')], vals[(9, 'host')])}", flush=True)
be.set_option("impl", 0)
for (N, d, kernel) in ((32769, 1024, "polynomial"), (32769, 4096, "rbf")):
    X, y = make_data(N, d, 9, np.float32)
    ds = be.dataset(X)
    q, k_last = be.run_q_kernel(ds, kernel)
    v = np.ones(N - 1, np.float32)
    for impl in (2, 6, 8, 9):
        be.set_option("impl", impl)
        ts = []
        for _ in range(4):
            out = be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
            t = be.timings()
            ts.append(t["matvec_tile_ms"])
        print(f"fp32 {N}x{d} {kernel} impl {impl}: tile kernel {min(ts):.3f} ms -> {t['matvec_flops'] / min(ts) / 1e9:.1f} TFLOP/s  all: {[round(x, 3) for x in ts]}", flush=True)
    be.set_option("impl", 0)
    del ds
