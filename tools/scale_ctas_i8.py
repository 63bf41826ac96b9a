"""Per-CTA throughput of the int8-slice tile kernel against the number of persistent CTAs (option "max_ctas"): separates the per-SM limit of the
kernel from chip-level contention (L2 -> SM bandwidth, power).  Run on the GPU box:  python tools/scale_ctas_i8.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import plssvm_b200 as pb
from datagen import make_data
be = pb.Backend(0)
for dtype, N, d, kernel in ((np.float64, 16385, 4096, "rbf"), (np.float32, 32769, 4096, "rbf"), (np.float32, 32769, 1024, "polynomial")):
    X, y = make_data(N, d, 9, dtype)
    ds = be.dataset(X)
    q, k_last = be.run_q_kernel(ds, kernel)
    v = np.ones(N - 1, dtype)
    be.set_option("impl", 6)
    for ctas in (148, 111, 74, 37):
        be.set_option("max_ctas", ctas)
        ts = []
        for _ in range(3):
            be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
            t = be.timings(); ts.append(t["matvec_tile_ms"])
        tf = t["matvec_flops"] / min(ts) / 1e9
        print(f"{np.dtype(dtype).name} {N}x{d} ctas {ctas}: {min(ts):.3f} ms  {tf:.1f} TFLOP/s  per-CTA {tf/ctas:.3f}", flush=True)
    be.set_option("max_ctas", 0)
    del ds
