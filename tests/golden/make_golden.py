#!/usr/bin/env python
"""Regenerates the committed golden fixtures.  Needs /root/reference (so it only runs in the build container).

Outputs (all under tests/golden/):

* ``predict_500x200.npz`` — the reference's own known-answer fixture for the prediction path
  (reference tests/backends/generic_csvm_tests.hpp:197-247): ``tests/data/predict/500x200_test.libsvm`` evaluated with
  ``500x200_{linear,polynomial,rbf}.libsvm.model`` must give exactly ``500x200.libsvm.predict``.
  The text files are parsed into float64 arrays (LIBSVM format: libsvm_parsing.hpp:117-221, model header
  libsvm_model_parsing.hpp:82-272); no reference source is copied.
* ``kernels_500x200.npz`` — the reference's property-test inputs (generic_csvm_tests.hpp:372-493:
  ``tests/data/libsvm/500x200.libsvm``, params {degree 2, gamma 1e-3, coef0 1, cost 0.1}) together with the outputs of the
  reference's OWN compiled OpenMP kernels (oracle/_ref/liboracle_ref.so): q, matvec (add = +1 / -1), w.
* ``ref_vectors.npz`` — outputs of the reference build on seeded synthetic inputs (tests/datagen.py), fp64 and fp32:
  q, matvec, full CG solve (alpha, rho, iterations, residual trace), predict values.  Inputs are regenerated from the seed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from datagen import GOLDEN_CASES, make_case  # noqa: E402

REF = "/root/reference"


def parse_libsvm(path, num_features=None, skip_until=None):
    """Dense parse of a LIBSVM file: returns (first column, X)."""
    first, rows = [], []
    started = skip_until is None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not started:
                started = line == skip_until
                continue
            if not line or line.startswith("#"):
                continue
            parts = line.split()
            first.append(float(parts[0]))
            rows.append({int(k): float(v) for k, v in (p.split(":") for p in parts[1:])})
    d = num_features or max(max(r) for r in rows if r)
    X = np.zeros((len(rows), d), dtype=np.float64)
    for i, r in enumerate(rows):
        for k, v in r.items():
            X[i, k - 1] = v
    return np.array(first, dtype=np.float64), X


def parse_model(path):
    hdr = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line == "SV":
                break
            k, *v = line.split()
            hdr[k] = v
    alpha, sv = parse_libsvm(path, skip_until="SV")
    return {
        "kernel": oracle.KERNEL_IDS[hdr["kernel_type"][0]],
        "degree": int(hdr.get("degree", ["3"])[0]),
        "gamma": float(hdr.get("gamma", ["0"])[0]),
        "coef0": float(hdr.get("coef0", ["0"])[0]),
        "rho": float(hdr["rho"][0]),
        "labels": [int(x) for x in hdr["label"]],
        "alpha": alpha,
        "sv": sv,
    }


def main():
    ref = oracle.Oracle("reference")
    assert ref.reported_kind() == "reference"

    # ---- 1. predict fixture -----------------------------------------------------------------------------------------
    labels, pts = parse_libsvm(f"{REF}/tests/data/predict/500x200_test.libsvm", 200)
    expected = np.loadtxt(f"{REF}/tests/data/predict/500x200.libsvm.predict").astype(np.int32)
    out = {"points": pts, "expected_labels": expected, "test_file_labels": labels.astype(np.int32)}
    for name in ("linear", "polynomial", "rbf"):
        m = parse_model(f"{REF}/tests/data/predict/500x200_{name}.libsvm.model")
        assert m["sv"].shape[1] == 200, m["sv"].shape
        for k in ("kernel", "degree", "gamma", "coef0", "rho", "alpha", "sv"):
            out[f"{name}_{k}"] = np.asarray(m[k])
        vals, _ = ref.predict(m["kernel"], m["sv"], m["alpha"], m["rho"], pts, m["degree"], m["gamma"], m["coef0"])
        got = oracle.sign_labels(vals)
        assert (got == expected).all(), f"{name}: reference build disagrees with the reference's own fixture"
        out[f"{name}_ref_values"] = vals
        print(f"predict/{name}: {len(m['alpha'])} SV, 500/500 labels reproduced, min |value| = {np.abs(vals).min():.3e}")
    np.savez_compressed(os.path.join(HERE, "predict_500x200.npz"), **out)

    # ---- 2. kernel property-test fixture -----------------------------------------------------------------------------
    y, X = parse_libsvm(f"{REF}/tests/data/libsvm/500x200.libsvm", 200)
    n = X.shape[0] - 1
    rng = np.random.Generator(np.random.Philox(20221017))
    rhs = rng.uniform(1.0, 2.0, n)          # generic_csvm_tests.hpp:456-458 — random rhs in [1, 2)
    alpha = rng.uniform(0.0, 1.0, X.shape[0])  # :415-417 — random weights in [0, 1)
    degree, gamma, coef0, cost = 2, 1e-3, 1.0, 0.1
    out = {"X": X, "y": y.astype(np.int32), "rhs": rhs, "alpha": alpha, "degree": degree, "gamma": gamma, "coef0": coef0, "cost": cost}
    for name, kid in oracle.KERNEL_IDS.items():
        q = ref.q(kid, X, degree, gamma, coef0)
        qa = ref.kernel_function(kid, X[-1], X[-1], degree, gamma, coef0) + 1.0 / cost
        out[f"{name}_q"] = q
        out[f"{name}_QA_cost"] = qa
        for add in (1.0, -1.0):
            out[f"{name}_matvec_{'p' if add > 0 else 'm'}"] = ref.matvec(kid, X, q, rhs, np.zeros(n), qa, 1.0 / cost, add, degree, gamma, coef0)
    out["w"] = ref.w(X, alpha)
    np.savez_compressed(os.path.join(HERE, "kernels_500x200.npz"), **out)
    print("kernels_500x200.npz written")

    # ---- 3. seeded synthetic cases ------------------------------------------------------------------------------------
    out = {}
    for case in GOLDEN_CASES:
        c = make_case(case)
        X, y, kid = c["X"], c["y"], c["kernel"]
        pr = dict(degree=c["degree"], gamma=c["gamma"], coef0=c["coef0"])
        name = c["name"]
        q = ref.q(kid, X, **pr)
        qa = ref.kernel_function(kid, X[-1], X[-1], **pr) + 1.0 / c["cost"]
        out[f"{name}/q"] = q
        out[f"{name}/QA_cost"] = np.asarray(qa, dtype=X.dtype)
        out[f"{name}/matvec_p"] = ref.matvec(kid, X, q, c["v"], np.zeros(X.shape[0] - 1, X.dtype), qa, 1.0 / c["cost"], 1.0, **pr)
        out[f"{name}/matvec_m"] = ref.matvec(kid, X, q, c["v"], c["v"], qa, 1.0 / c["cost"], -1.0, **pr)
        res = ref.solve(kid, X, y, cost=c["cost"], eps=c["eps"], max_iter=c["max_iter"], trace=True, **pr)
        # The reference is not run-to-run reproducible (atomics: SURVEY.md §5) and CG amplifies the rounding noise, so the
        # fixture also records the reference's OWN spread over repeated runs with 8/1/3/5 threads: parity tolerances for
        # alpha/rho are max(stated tolerance, 20 x this spread).
        spread_a, spread_r, iters = 0.0, 0.0, {res["iterations"]}
        for nthr in (8, 1, 3, 5, 8):
            ref.set_threads(nthr)
            again = ref.solve(kid, X, y, cost=c["cost"], eps=c["eps"], max_iter=c["max_iter"], **pr)
            iters.add(again["iterations"])
            if again["iterations"] == res["iterations"]:
                spread_a = max(spread_a, float(np.max(np.abs(again["alpha"] - res["alpha"])) / np.max(np.abs(res["alpha"]))))
                spread_r = max(spread_r, float(abs(again["rho"] - res["rho"])))
        ref.set_threads(8)
        out[f"{name}/alpha_spread"] = np.asarray(spread_a)
        out[f"{name}/rho_spread"] = np.asarray(spread_r)
        out[f"{name}/iterations_seen"] = np.asarray(sorted(iters))
        out[f"{name}/alpha"] = res["alpha"]
        out[f"{name}/rho"] = np.asarray(res["rho"])
        out[f"{name}/iterations"] = np.asarray(res["iterations"])
        out[f"{name}/trace"] = res["trace"]
        vals, w = ref.predict(kid, X, res["alpha"], res["rho"], c["P"], **pr)
        out[f"{name}/predict"] = vals
        if w is not None:
            out[f"{name}/w"] = w
        print(f"{name}: spread alpha {spread_a:.1e} rho {spread_r:.1e} iters {sorted(iters)}")
        print(f"{name}: {res['iterations']} iterations, delta {res['delta']:.3e} / delta0 {res['delta0']:.3e}, min|f| {np.abs(vals).min():.2e}")
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)


if __name__ == "__main__":
    main()
