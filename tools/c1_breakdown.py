#!/usr/bin/env python
"""Where the time of a small whole fit (C1: 5,000 x 1,000 linear fp64 from host buffers) goes: wall clock of plssvm_b200_solve vs the library's own timers."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plssvm_b200 as pb  # noqa: E402
from bench import WORKLOADS, make_host_data  # noqa: E402

N, d, kernel, dtype, _ = WORKLOADS["C1"]
X, y = make_host_data(N, d, dtype, 42)
be = pb.Backend(0)
for rep in range(4):
    t0 = time.perf_counter()
    r = be.solve(X, y, kernel, eps=1e-8)
    wall = (time.perf_counter() - t0) * 1e3
    t = be.timings()
    print(f"rep {rep}: wall {wall:.2f} ms  total_ms {t['total_ms']:.2f}  cg_loop_ms {t['cg_loop_ms']:.2f}  matvec_ms {t['matvec_ms']:.2f}  tile_ms {t['matvec_tile_ms']:.2f}  "
          f"matvecs {t['matvec_calls']}  launches {t['kernel_launches']}  iterations {r['iterations']}  h2d {t['h2d_bytes'] / 1e6:.1f} MB", flush=True)
ds = be.dataset(X)
for rep in range(2):
    t0 = time.perf_counter()
    r = be.solve(ds, y, kernel, eps=1e-8)
    wall = (time.perf_counter() - t0) * 1e3
    t = be.timings()
    print(f"resident rep {rep}: wall {wall:.2f} ms  total_ms {t['total_ms']:.2f}  cg_loop_ms {t['cg_loop_ms']:.2f}  matvec_ms {t['matvec_ms']:.2f}  tile_ms {t['matvec_tile_ms']:.2f}", flush=True)
