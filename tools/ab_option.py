#!/usr/bin/env python
"""A/B of a tile-kernel option on one matvec: bit-identity of the result and tile-kernel time, per shape.
    python tools/ab_option.py i8_a_via_tmem [workload rows features]..."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plssvm_b200 as pb  # noqa: E402
from bench import WORKLOADS, make_device_data, matvec_flops  # noqa: E402

option = sys.argv[1]
shapes = [("C2", 1100, 200), ("C2", 16384, 4096), ("C2", 65536, 4096)] if len(sys.argv) < 3 else [(sys.argv[i], int(sys.argv[i + 1]), int(sys.argv[i + 2])) for i in range(2, len(sys.argv), 3)]
be = pb.Backend(0)
dev = torch.device("cuda", 0)
for workload, rows, feats in shapes:
    _, _, kernel, dtype, _ = WORKLOADS[workload]
    X, _ = make_device_data(rows, feats, dtype, 7, dev)
    ds = be.dataset(X)
    q, k_last = be.run_q_kernel(ds, kernel)
    v = np.random.default_rng(1).uniform(1, 2, rows - 1).astype(np.dtype(dtype))
    out = {"option": option, "workload": workload, "rows": rows, "features": feats}
    res = {}
    for val in (0, 1, 0, 1):
        be.set_option(option, val)
        ms = []
        for _ in range(3):
            res[val] = be.run_svm_kernel(ds, q, v, np.zeros_like(v), float(k_last) + 1.0, 1.0, 1.0, kernel)
            ms.append(be.timings()["matvec_tile_ms"])
        out.setdefault(f"tflops_{val}", []).append(round(matvec_flops(rows, feats) / (min(ms) * 1e-3) / 1e12, 2))
    be.set_option(option, 0)
    out["bit_identical"] = bool(np.array_equal(res[0], res[1]))
    out["max_rel_diff"] = float(np.max(np.abs(res[0] - res[1])) / np.max(np.abs(res[0])))
    print(json.dumps(out), flush=True)
    ds.close()
    del X
    torch.cuda.empty_cache()
