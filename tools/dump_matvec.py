#!/usr/bin/env python
"""One implicit matvec of a bench workload shape with the library named by PLSSVM_B200_LIB, result saved as .npy — for bit-identity checks ACROSS two builds
of the library (A/B builds of a kernel change):  python tools/dump_matvec.py out.npy C2 2049 333"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plssvm_b200 as pb  # noqa: E402
from bench import WORKLOADS, make_device_data  # noqa: E402

out, workload, rows, feats = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
_, _, kernel, dtype, _ = WORKLOADS[workload]
be = pb.Backend(0)
X, _ = make_device_data(rows, feats, dtype, 7, torch.device("cuda", 0))
ds = be.dataset(X)
q, k_last = be.run_q_kernel(ds, kernel)
v = np.random.default_rng(1).uniform(1, 2, rows - 1).astype(np.dtype(dtype))
res = be.run_svm_kernel(ds, q, v, np.zeros_like(v), float(k_last) + 1.0, 1.0, 1.0, kernel)
np.save(out, res)
print(out, be.timings()["impl_used"], float(np.abs(res).max()))
