"""CPU tests pinning the extended-precision oracle (oracle/lssvm_exact.cpp), the noise-free parity target.

It must (1) reproduce the reference's known-answer tests, (2) agree with the reference-compiled build to the reference's own rounding
level on every golden case, (3) be deterministic — independent of the thread count and identical to the committed fixture
tests/golden/exact_vectors.npz bit for bit (the GPU tests compare against that fixture, not against a live run).
"""
import os

import numpy as np
import pytest

import oracle
from datagen import GOLDEN_CASES, make_case


@pytest.fixture(scope="module")
def ex():
    return oracle.Exact()


def test_trivial_solve_and_predict(ex):
    """generic_csvm_tests.hpp:99-137 / 149-195: A = sqrt(1 - 1/C) I, rhs (1, -1, 1, -1) -> x = rhs, rho = 0;  predict -> {0, 4}."""
    C = 2.0
    A = np.sqrt(1.0 - 1.0 / C) * np.eye(4)
    rhs = np.array([1.0, -1.0, 1.0, -1.0])
    r = ex.solve(oracle.LINEAR, A, rhs, cost=C, eps=1e-10)
    assert np.allclose(r["alpha"], rhs, atol=1e-14) and abs(r["rho"]) < 1e-14
    sv = np.eye(4)
    vals = ex.predict(oracle.LINEAR, sv, np.ones(4), 0.0, np.array([[0.0] * 4, [1.0] * 4]))
    assert np.allclose(vals, [0.0, 4.0], atol=1e-15)


@pytest.mark.parametrize("kernel", ["linear", "polynomial", "rbf"])
def test_predict_fixture_labels(ex, kernel, golden_dir):
    """The reference's 500x200 prediction fixture (generic_csvm_tests.hpp:197-247): exact labels."""
    g = np.load(os.path.join(golden_dir, "predict_500x200.npz"))
    vals = ex.predict(int(g[f"{kernel}_kernel"]), g[f"{kernel}_sv"], g[f"{kernel}_alpha"], float(g[f"{kernel}_rho"]), g["points"], int(g[f"{kernel}_degree"]),
                      float(g[f"{kernel}_gamma"]), float(g[f"{kernel}_coef0"]))
    assert np.array_equal(oracle.sign_labels(vals), g["expected_labels"])
    assert np.max(np.abs(vals - g[f"{kernel}_ref_values"])) <= 1e-12 * np.max(np.abs(vals))


@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c["name"] for c in GOLDEN_CASES])
def test_agrees_with_the_reference_build_and_the_fixture(ex, case, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_vectors.npz"))
    e = np.load(os.path.join(golden_dir, "exact_vectors.npz"))
    c = make_case(case)
    X, y, kid, name = c["X"], c["y"], c["kernel"], c["name"]
    pr = dict(degree=c["degree"], gamma=c["gamma"], coef0=c["coef0"])
    f64 = X.dtype == np.float64
    # one application of the kernels: the reference's rounding level is a few n * eps of its real type
    mv = ex.matvec(kid, X, g[f"{name}/q"], c["v"], float(g[f"{name}/QA_cost"]), 1.0 / c["cost"], **pr)
    assert np.array_equal(mv, e[f"{name}/matvec"]), "exact matvec differs from the committed fixture"
    assert float(e[f"{name}/ref_matvec_err"]) <= (1e-14 if f64 else 5e-6)
    q = ex.q(kid, X, **pr)
    assert np.max(np.abs(q[:-1] - g[f"{name}/q"])) <= (1e-14 if f64 else 5e-6) * max(1.0, float(np.max(np.abs(q))))
    # CG after exactly k iterations is reproduced bit for bit with another thread count
    k = int(e[f"{name}/iterations"])
    os.environ["OMP_NUM_THREADS"] = "3"
    r = ex.solve(kid, X, y, cost=c["cost"], eps=1e-30 if f64 else 1e-18, max_iter=k, **pr)
    assert r["iterations"] == k and np.array_equal(r["alpha"], e[f"{name}/alpha_k"]) and r["rho"] == float(e[f"{name}/rho_k"])
    # the reference's residual history follows the exact one while rounding noise is not yet amplified
    m = min(3, len(r["trace"]), len(g[f"{name}/trace"]))
    assert np.allclose(g[f"{name}/trace"][:m], r["trace"][:m], rtol=1e-8 if f64 else 2e-2)
    # the converged solution solves the reduced system: alpha sums to zero and the exact matvec reproduces b~ = y - y_N
    star = e[f"{name}/alpha_star"]
    assert abs(star.sum()) <= 1e-12 * np.abs(star).sum()
    if X.shape[0] > 2:
        qe = ex.q(kid, X, **pr)
        Ax = ex.matvec(kid, X.astype(np.float64), qe[:-1], star[:-1], qe[-1] + 1.0 / c["cost"], 1.0 / c["cost"], **pr)
        b = (y[:-1] - y[-1]).astype(np.float64)
        assert np.max(np.abs(Ax - b)) <= 1e-9 * max(1.0, float(np.max(np.abs(star))) * X.shape[0]), "x* does not solve the reduced system"
