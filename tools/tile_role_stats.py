#!/usr/bin/env python
"""Where the warp-specialised roles of the int8-slice tile kernel wait (option "tile_stats"): one matvec per shape, printed as JSON lines.
    python tools/tile_role_stats.py            # C2 (fp64 rbf 65536 x 4096), C3 (fp32 poly 131072 x 1024) and smaller slices of both
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plssvm_b200 as pb  # noqa: E402
from bench import WORKLOADS, make_device_data, matvec_flops  # noqa: E402

be = pb.Backend(0)
dev = torch.device("cuda", 0)
for workload, rows, feats in (("C2", 65536, 4096), ("C2", 16384, 4096), ("C2", 16384, 1024), ("C3", 131072, 1024), ("C3", 32768, 1024), ("C3", 32768, 4096)):
    _, _, kernel, dtype, _ = WORKLOADS[workload]
    X, _ = make_device_data(rows, feats, dtype, 7, dev)
    ds = be.dataset(X)
    q, k_last = be.run_q_kernel(ds, kernel)
    v = np.ones(rows - 1, dtype=np.dtype(dtype))
    out = {"workload": workload, "rows": rows, "features": feats, "kernel": kernel, "dtype": dtype}
    for stats in (0, 1):
        be.set_option("tile_stats", stats)
        for _ in range(3):
            be.run_svm_kernel(ds, q, v, np.zeros_like(v), float(k_last) + 1.0, 1.0, 1.0, kernel)
        t = be.timings()
        if stats == 0:
            out["tile_ms"] = t["matvec_tile_ms"]
            out["tflops"] = matvec_flops(rows, feats) / (t["matvec_tile_ms"] * 1e-3) / 1e12
        else:
            out.update({k: round(t[k], 4) for k in ("tile_mma_wait_operands", "tile_mma_wait_drain", "tile_producer_wait", "tile_epilogue_wait")})
    be.set_option("tile_stats", 0)
    print(json.dumps(out), flush=True)
    ds.close()
    del X
    torch.cuda.empty_cache()
