"""Run under torchrun (one rank per GPU): the tile-sharded, NCCL-reduced solve must reproduce the single-GPU solve and the
oracle.  Used by tests/test_gpu_multi.py and runnable by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import plssvm_b200 as pb  # noqa: E402
from datagen import make_data  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = pb.Backend(local)
    be.init_comm_from_torch()
    single = pb.Backend(local)  # same GPU, no communicator: the unsharded result
    orc = oracle.Oracle("reference" if oracle.available("reference") else "port")
    failures = []
    for dtype, kernel, kid, N, d, eps in ((np.float64, "rbf", 2, 1500, 100, 1e-8), (np.float64, "linear", 0, 1100, 64, 1e-8), (np.float32, "polynomial", 1, 900, 48, 1e-4)):
        X, y = make_data(N, d, 500 + kid, dtype)
        n = N - 1
        ds, ds1 = be.dataset(X), single.dataset(X)
        q, k_last = be.run_q_kernel(ds, kernel)
        v = np.random.default_rng(4).uniform(1, 2, n).astype(dtype)
        got = be.run_svm_kernel(ds, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
        one = single.run_svm_kernel(ds1, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, kernel)
        want = orc.matvec(kid, X, q, v, np.zeros_like(v), k_last + 1.0, 1.0, 1.0, gamma=1.0 / d)
        tol = 1e-12 if dtype == np.float64 else 2e-4
        scale = float(np.max(np.abs(want)))
        for name, a, b in (("sharded vs oracle", got, want), ("sharded vs single", got, one)):
            err = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) / scale
            if not err < tol:
                failures.append(f"{kernel}/{np.dtype(dtype).name} matvec {name}: {err:.3e}")
        # every rank must hold the identical result (the vector updates run redundantly on each rank)
        t = torch.from_numpy(got.astype(np.float64)).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            failures.append(f"{kernel}: ranks disagree on the reduced matvec")
        r = be.solve(ds, y, kernel, eps=eps)
        r1 = single.solve(ds1, y, kernel, eps=eps)
        if abs(r["iterations"] - r1["iterations"]) > 1:
            failures.append(f"{kernel}: iterations sharded {r['iterations']} vs single {r1['iterations']}")
        elif r["iterations"] == r1["iterations"] and dtype == np.float64:
            # the all-reduce changes the summation order; CG amplifies that to the noise floor of this data family.
            # fp32 solves here are decided by rounding noise within 2-3 iterations, so only fp64 is compared element-wise.
            err = float(np.max(np.abs(r["alpha"] - r1["alpha"])) / np.max(np.abs(r1["alpha"])))
            if not err < 1e-5:
                failures.append(f"{kernel}: alpha sharded vs single {err:.3e}")
        for res in (r, r1):
            if not res["delta"] <= eps * eps * res["delta0"]:
                failures.append(f"{kernel}: stopping criterion not met: {res['delta']} > {eps * eps * res['delta0']}")
        a = torch.from_numpy(r["alpha"].astype(np.float64)).cuda()
        lo, hi = a.clone(), a.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            failures.append(f"{kernel}: ranks disagree on alpha")
        # predict: every rank gets the same points, computes its range, and the ranges are exchanged — every rank returns ALL values, equal to the unsharded ones
        P, _ = make_data(1000, d, 700 + kid, dtype)
        vals, _ = be.predict_values(X, r1["alpha"], r1["rho"], P, kernel)
        vals1, _ = single.predict_values(X, r1["alpha"], r1["rho"], P, kernel)
        if not np.array_equal(vals, vals1):
            failures.append(f"{kernel}: sharded predict differs from the unsharded one by {np.max(np.abs(vals - vals1)):.3e}")
        if rank == 0:
            print(f"[world {world}] {kernel}/{np.dtype(dtype).name}: iterations {r['iterations']} (single {r1['iterations']})", flush=True)
    dist.barrier()
    if failures:
        print(f"rank {rank} FAILURES:\n  " + "\n  ".join(failures), flush=True)
    elif rank == 0:
        print("multi-gpu check passed", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
