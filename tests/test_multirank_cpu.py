"""world_size-2 tests on CPU (gloo): the host-side logic of the multi-GPU path — tile sharding by rank through the
library's own schedule functions, the sum of the per-rank partial results (the NCCL all-reduce on the GPU), and the
unique-id broadcast plumbing.  The per-rank partial matvec is emulated with numpy on the dense Q~ (built from the oracle's
kernel function semantics); the GPU kernels themselves are covered by the `-m gpu` tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import plssvm_b200 as pb
from datagen import make_data


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dense_Q(X, q, qa, cost_inv, kernel, gamma):
    A = X[:-1]
    G = A @ A.T
    if kernel == 0:
        K = G
    elif kernel == 1:
        K = (gamma * G) ** 3
    else:
        sq = np.sum(A * A, axis=1)
        K = np.exp(-gamma * np.maximum(sq[:, None] + sq[None, :] - 2 * G, 0))
    return K + qa - q[:, None] - q[None, :] + cost_inv * np.eye(len(q))


def _worker(rank, world, port, kernel, out_queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the id broadcast used by Backend.init_comm_from_torch
        payload = bytes(range(128)) if rank == 0 else b""
        assert pb.broadcast_bytes(payload, 128) == bytes(range(128))

        # 2. sharded matvec: each rank sums only the tiles it owns, then all-reduce
        X, _ = make_data(700, 24, 321, np.float64)
        n = X.shape[0] - 1
        orc = oracle.Oracle("port")
        gamma = 1.0 / 24
        q = orc.q(kernel, X, gamma=gamma)
        qa = orc.kernel_function(kernel, X[-1], X[-1], gamma=gamma) + 1.0
        v = np.random.default_rng(9).uniform(1, 2, n)
        Q = _dense_Q(X, q, qa, 1.0, kernel, gamma)
        tile = pb.tile_size()
        T = (n + tile - 1) // tile
        lo, hi = pb.rank_range(pb.tri_num_tiles(T), rank, world)
        part = np.zeros(n)
        for L in range(lo, hi):
            I, J = pb.tri_decode(T, L)
            r = slice(I * tile, min(n, (I + 1) * tile))
            c = slice(J * tile, min(n, (J + 1) * tile))
            part[r] += Q[r, c] @ v[c]
            if I != J:  # mirrored contribution of an off-diagonal tile
                part[c] += Q[r, c].T @ v[r]
        t = torch.from_numpy(part)
        dist.all_reduce(t)
        want = orc.matvec(kernel, X, q, v, np.zeros(n), qa, 1.0, 1.0, gamma=gamma)
        err = float(np.max(np.abs(t.numpy() - want)) / np.max(np.abs(want)))
        out_queue.put((rank, hi - lo, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kernel", [0, 2])
def test_two_rank_sharded_matvec_matches_oracle(kernel):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kernel, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    counts = [c for _, c, _ in res]
    assert abs(counts[0] - counts[1]) <= 1 and sum(counts) == pb.tri_num_tiles((699 + 127) // 128)
    for _, _, err in res:
        assert err < 1e-12


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_weighted_ranges_tile_the_order_exactly(world):
    """Rate-weighted shares (option "balance"): whatever the weights, the ranks' ranges are contiguous, ordered and cover every tile exactly once;
    equal weights reproduce rank_range; degenerate weights fall back to equal shares."""
    rng = np.random.default_rng(world)
    for total in (0, 1, 7, 131328, 10 ** 9 + 7):
        for weights in ([1.0] * world, list(rng.uniform(0.75, 1.25, world)), list(rng.uniform(0.01, 5.0, world))):
            edges = [pb.weighted_range(total, g, world, weights) for g in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(lo <= hi for lo, hi in edges) and all(edges[g][1] == edges[g + 1][0] for g in range(world - 1))
            if len(set(weights)) == 1:
                assert edges == [pb.rank_range(total, g, world) for g in range(world)]
            elif total > 1000 * world:
                shares = np.array([hi - lo for lo, hi in edges], dtype=float) / total
                assert np.allclose(shares, np.array(weights) / np.sum(weights), atol=2.0 / total + 1e-12)
        bad = [1.0] * world
        bad[-1] = 0.0
        assert [pb.weighted_range(total, g, world, bad) for g in range(world)] == [pb.rank_range(total, g, world) for g in range(world)]
