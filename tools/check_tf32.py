"""Quick hardware check of the fp32 tensor path (tcgen05 3xTF32) against the SIMT tiles and the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import plssvm_b200 as pb  # noqa: E402
from datagen import make_data  # noqa: E402

be = pb.Backend(0)
orc = oracle.Oracle("reference" if oracle.available("reference") else "port")
for (N, d) in ((130, 40), (300, 64), (386, 64), (1000, 96), (2049, 333)):
    for kernel, kid in (("linear", 0), ("polynomial", 1), ("rbf", 2)):
        X, y = make_data(N, d, 7, np.float32)
        n = N - 1
        ds = be.dataset(X)
        q, k_last = be.run_q_kernel(ds, kernel)
        v = np.random.default_rng(3).uniform(1, 2, n).astype(np.float32)
        outs = {}
        for impl in (1, 2, 4, 5):
            be.set_option("impl", impl)
            outs[impl] = be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
        be.set_option("impl", 0)
        X64 = X.astype(np.float64)
        want = orc.matvec(kid, X64, q.astype(np.float64), v.astype(np.float64), np.zeros(n), float(k_last) + 1.0, 1.0, 1.0, gamma=1.0 / d)
        sc = np.max(np.abs(want))
        e1 = np.max(np.abs(outs[1] - want)) / sc
        e2 = np.max(np.abs(outs[2] - want)) / sc
        e4 = np.max(np.abs(outs[4] - want)) / sc
        e5 = np.max(np.abs(outs[5] - want)) / sc
        print(f"N={N:5d} d={d:4d} {kernel:10s} simt {e1:.2e}  tf32x3 {e2:.2e}  CTA-pair {e4:.2e}  128x256 {e5:.2e}  (vs fp64 oracle)", flush=True)
# timing on a bigger problem
X, y = make_data(32769, 1024, 9, np.float32)
ds = be.dataset(X)
q, k_last = be.run_q_kernel(ds, "polynomial")
v = np.ones(32768, np.float32)
for impl in (1, 2, 4, 5):
    be.set_option("impl", impl)
    for _ in range(3):
        be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, "polynomial")
        t = be.timings()
    print(f"impl {impl}: tile kernel {t['matvec_tile_ms']:.3f} ms -> {t['matvec_flops'] / t['matvec_tile_ms'] / 1e9:.1f} TFLOP/s", flush=True)

# predict (rectangular tiles) with the CTA-pair kernel vs the single-CTA kernel
P, _ = make_data(700, 1024, 10, np.float32)
alpha = np.random.default_rng(5).standard_normal(32769).astype(np.float32)
vals = {}
for impl in (2, 4, 5):
    be.set_option("impl", impl)
    vals[impl], _ = be.predict_values(ds, alpha, 0.1, be.dataset(P), "rbf")
for impl in (4, 5):
    print(f"predict rbf: max |impl{impl} - impl2| / scale =", float(np.max(np.abs(vals[impl] - vals[2])) / np.max(np.abs(vals[2]))), flush=True)
be.set_option("impl", 0)
