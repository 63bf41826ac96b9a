"""Small end-to-end exercise of every kernel of libplssvm_b200.so, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck) on the GPU box:  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import plssvm_b200 as pb  # noqa: E402
from datagen import make_data  # noqa: E402

be = pb.Backend(0)
# the default library holds impl 1 / 2 / 6 / 7 and, for fp32, 10 (CTA pairs); the variants 4 / 5 / 8 / 9 and the fp64 CTA-pair kernel only exist in a build
# with -DPLSSVM_B200_EXPERIMENTAL
impls = {np.float64: (1, 2, 6, 8, 10), np.float32: (1, 2, 4, 5, 6, 7, 8, 9, 10)} if pb.has_experimental() else {np.float64: (1, 2, 6), np.float32: (1, 2, 6, 7, 10)}
for dtype in (np.float64, np.float32):
    X, y = make_data(301, 37, 1, dtype)
    P, _ = make_data(150, 37, 2, dtype)
    for kernel in ("linear", "polynomial", "rbf"):
        for impl in impls[dtype]:
            be.set_option("impl", impl)
            r = be.solve(X, y, kernel, eps=1e-6 if dtype == np.float64 else 1e-3, max_iter=60)
            vals, _ = be.predict_values(X, r["alpha"], r["rho"], P, kernel)
            assert np.all(np.isfinite(vals)), (dtype, kernel, impl)
        be.set_option("impl", 0)
    be.set_option("linear_factorized", 1)
    be.solve(X, y, "linear", eps=1e-6 if dtype == np.float64 else 1e-3)
    be.set_option("linear_factorized", 0)
    ds = be.dataset(X)
    be.run_w_kernel(ds, np.ones(301, dtype))
    # round 2: row-pointer entry points (pinned ring), badly scaled predict batch (per-batch guard + fallback), A planes through tensor memory,
    # the slow-drain A/B path, per-role counters, virtual ranks (sharded matvec / predict on one device)
    r = be.solve_rows(X, y, "rbf", eps=1e-6 if dtype == np.float64 else 1e-3, max_iter=40)
    Pb = P.copy()
    Pb[3, 0] *= dtype(2.0 ** 26)
    be.predict_values_rows(X, r["alpha"], r["rho"], Pb, "rbf")
    q, k_last = be.run_q_kernel(ds, "rbf")
    v = np.ones(300, dtype)
    for opt in ("i8_a_via_tmem", "tile_stats"):
        be.set_option(opt, 1)
        be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, "rbf")
        be.set_option(opt, 0)
    be.set_option("fp32_fast_drain", 0)
    be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, "polynomial")
    be.set_option("fp32_fast_drain", 1)
    be.set_option("virtual_world", 3)
    for g in range(3):
        be.set_option("virtual_rank", g)
        be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, "rbf")
        be.predict_values(X, r["alpha"], r["rho"], P, "polynomial")
    be.set_option("virtual_world", 1)
    ds.close()
print("sanitize_run finished")
