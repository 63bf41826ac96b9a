"""Pins the oracle (CPU, no GPU needed): both builds against the reference's known-answer tests and committed golden
vectors, and the restated `port` build against the reference-compiled build.

Tolerance mirrors the reference's EXPECT_FLOATING_POINT_NEAR (tests/custom_test_macros.hpp:114-137):
|a - b| < max(min_normal, 128 * eps_machine * (|a| + |b|)).
"""
import os

import numpy as np
import pytest

import oracle
from datagen import GOLDEN_CASES, make_case
from parity import check_solution

KERNELS = {"linear": 0, "polynomial": 1, "rbf": 2}


def ref_near(a, b, factor=128.0):
    a = np.asarray(a)
    b = np.asarray(b)
    fi = np.finfo(a.dtype)
    tol = np.maximum(fi.tiny, factor * fi.eps * (np.abs(a) + np.abs(b)))
    return np.all(np.abs(a - b) < tol)


@pytest.fixture(scope="module", params=["port", "reference"])
def orc(request):
    if not oracle.available(request.param) and request.param != "port":
        pytest.skip("reference build not present")
    o = oracle.Oracle(request.param)
    assert o.reported_kind() == request.param
    return o


# ---- reference known-answer tests --------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["linear", "polynomial"])  # rbf is GTEST_SKIPped in the reference (generic_csvm_tests.hpp:111)
def test_solve_trivial(orc, kernel, dtype):
    """generic_csvm_tests.hpp:99-137: A = sqrt(1 - 1/C) I_4, C = 2 -> x == rhs, |rho| ~ 0."""
    C = 2.0
    A = np.eye(4, dtype=dtype) * dtype(np.sqrt(dtype(1.0) - dtype(1.0 / C)))
    rhs = np.array([1, -1, 1, -1], dtype=dtype)
    res = orc.solve(KERNELS[kernel], A, rhs, degree=1, gamma=1.0, coef0=0.0, cost=C, eps=1e-5, max_iter=4)
    assert ref_near(res["alpha"], rhs)
    assert abs(res["rho"]) < 4 * np.finfo(dtype).eps


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["linear", "polynomial"])
def test_predict_values_trivial(orc, kernel, dtype):
    """generic_csvm_tests.hpp:149-195: unit SVs, weights +-1 -> {0, 4}; linear fills w == weights, others leave it empty."""
    sv = np.eye(4, dtype=dtype)
    weights = np.array([1, -1, 1, -1], dtype=dtype)
    pts = np.array([[1, 1, 1, 1], [1, -1, 1, -1]], dtype=dtype)
    vals, w = orc.predict(KERNELS[kernel], sv, weights, 0.0, pts, degree=1, gamma=1.0, coef0=0.0)
    assert ref_near(vals, np.array([0, 4], dtype=dtype))
    if kernel == "linear":
        assert ref_near(w, weights)
    else:
        assert w is None


@pytest.mark.parametrize("kernel", ["linear", "polynomial", "rbf"])
def test_predict_fixture_labels(orc, kernel, golden_dir):
    """generic_csvm_tests.hpp:197-247: 500x200 test file x stored models -> exactly the stored labels, accuracy 1.0."""
    g = np.load(os.path.join(golden_dir, "predict_500x200.npz"))
    vals, _ = orc.predict(int(g[f"{kernel}_kernel"]), g[f"{kernel}_sv"], g[f"{kernel}_alpha"], float(g[f"{kernel}_rho"]), g["points"],
                          int(g[f"{kernel}_degree"]), float(g[f"{kernel}_gamma"]), float(g[f"{kernel}_coef0"]))
    labels = oracle.sign_labels(vals)
    assert (labels == g["expected_labels"]).all()
    assert (labels == g["test_file_labels"]).mean() == 1.0  # `score` test
    assert np.allclose(vals, g[f"{kernel}_ref_values"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("kernel", ["linear", "polynomial", "rbf"])
def test_kernel_fixture(orc, kernel, golden_dir):
    """generic_csvm_tests.hpp:372-493 inputs (500x200.libsvm, {deg 2, gamma 1e-3, coef0 1, C 0.1}) vs the reference kernels' outputs."""
    g = np.load(os.path.join(golden_dir, "kernels_500x200.npz"))
    X, k = g["X"], KERNELS[kernel]
    pr = dict(degree=int(g["degree"]), gamma=float(g["gamma"]), coef0=float(g["coef0"]))
    q = orc.q(k, X, **pr)
    assert ref_near(q, g[f"{kernel}_q"])
    qa = float(g[f"{kernel}_QA_cost"])
    assert ref_near(np.float64(orc.kernel_function(k, X[-1], X[-1], **pr) + 1.0 / float(g["cost"])), np.float64(qa))
    n = X.shape[0] - 1
    for add, tag in ((1.0, "p"), (-1.0, "m")):
        got = orc.matvec(k, X, q, g["rhs"], np.zeros(n), qa, 1.0 / float(g["cost"]), add, **pr)
        # atomics make the summation order run-dependent: compare with n * eps head-room, relative to the row scale
        want = g[f"{kernel}_matvec_{tag}"]
        assert np.max(np.abs(got - want)) <= 64 * np.finfo(np.float64).eps * np.max(np.abs(want)) * np.sqrt(n)
    if kernel == "linear":
        assert ref_near(orc.w(X, g["alpha"]), g["w"], factor=1e6)  # calculate_w test uses NEAR_EPS with factor 1e6 (:401-437)


def test_layout_contract():
    """generic_csvm_tests.hpp:560-593 pins the reference's SoA device layout {1,4,7,2,5,8,3,6,9}: our boundary takes the
    row-major host matrix instead — the oracle consumes the same row-major rows."""
    X = np.arange(1, 10, dtype=np.float64).reshape(3, 3)
    rows = [list(r) for r in X]
    assert rows == [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    assert list(X.T.reshape(-1)) == [1, 4, 7, 2, 5, 8, 3, 6, 9]


# ---- committed golden vectors from the reference build -------------------------------------------------------------------
@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c["name"] for c in GOLDEN_CASES])
def test_port_matches_reference_vectors(port, case, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_vectors.npz"))
    c = make_case(case)
    name, X, k = c["name"], c["X"], c["kernel"]
    fi = np.finfo(X.dtype)
    pr = dict(degree=c["degree"], gamma=c["gamma"], coef0=c["coef0"])
    q = port.q(k, X, **pr)
    assert ref_near(q, g[f"{name}/q"])
    qa = float(g[f"{name}/QA_cost"])
    n = X.shape[0] - 1
    mv = port.matvec(k, X, q, c["v"], np.zeros(n, X.dtype), qa, 1.0 / c["cost"], 1.0, **pr)
    want = g[f"{name}/matvec_p"]
    assert np.max(np.abs(mv - want)) <= 64 * fi.eps * max(np.max(np.abs(want)), 1e-30) * np.sqrt(max(n, 1))
    res = port.solve(k, X, c["y"], cost=c["cost"], eps=c["eps"], max_iter=c["max_iter"], trace=True, **pr)
    assert abs(res["iterations"] - int(g[f"{name}/iterations"])) <= 1
    # first residuals must agree tightly (before CG has amplified the rounding noise)
    assert np.allclose(res["trace"][:2], g[f"{name}/trace"][:2], rtol=1e-9 if X.dtype == np.float64 else 1e-3)
    if res["iterations"] == int(g[f"{name}/iterations"]):
        check_solution(res["alpha"], res["rho"], g[f"{name}/alpha"], float(g[f"{name}/rho"]), X.dtype, spread=float(g[f"{name}/alpha_spread"]), qa_cost=qa, tag=name)
    vals, _ = port.predict(k, X, g[f"{name}/alpha"], float(g[f"{name}/rho"]), c["P"], **pr)
    assert (oracle.sign_labels(vals) == oracle.sign_labels(g[f"{name}/predict"])).all()


def test_sign_zero_is_negative():
    """operators.hpp:178-181."""
    assert list(oracle.sign_labels(np.array([0.0, -0.0, 1e-300, -1e-300]))) == [-1, -1, 1, -1]
