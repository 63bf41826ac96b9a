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
impls = {np.float64: (1, 2, 6, 8), np.float32: (1, 2, 4, 5, 6, 7, 8, 9)}
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
print("sanitize_run finished")
