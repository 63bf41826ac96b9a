#!/usr/bin/env python
"""Regenerates tests/golden/exact_vectors.npz: the extended-precision targets (oracle/lssvm_exact.cpp) for the seeded golden cases of
tests/datagen.py, together with the errors of the REFERENCE build (oracle/_ref/liboracle_ref.so, the reference's own OpenMP kernels)
against them.  Needs /root/reference only for the reference-error entries (the exact targets themselves depend on nothing but the seed).

Per case `name`:
  name/matvec            exact Q~ v (add = +1, ret = 0) for the golden q / QA_cost (real-type values, the run_svm_kernel argument convention)
  name/ref_matvec_err    max |reference - exact| / max |exact|
  name/iterations        k = the reference's iteration count
  name/alpha_k, rho_k    exact CG after exactly k iterations (no stopping test)
  name/alpha_star, ...   the converged solution of the reduced system (exact CG to a relative residual of 1e-16)
  name/ref_alpha_err_k, ref_rho_err_k, ref_alpha_err_star, ref_rho_err_star
                         errors of the reference's solution (max over 8 / 1 / 3 / 5 OpenMP threads and the committed golden run),
                         alpha relative to max |alpha|, rho absolute
  name/predict           exact decision values of the golden model (the reference's alpha / rho as stored) on the case's test points
  name/ref_predict_err   max |reference values - exact| / max |exact|
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


def main():
    ex = oracle.Exact()
    ref = oracle.Oracle("reference")
    g = np.load(os.path.join(HERE, "ref_vectors.npz"))
    out = {}
    for case in GOLDEN_CASES:
        c = make_case(case)
        X, y, kid, name = c["X"], c["y"], c["kernel"], c["name"]
        pr = dict(degree=c["degree"], gamma=c["gamma"], coef0=c["coef0"])
        tiny = 1e-30 if X.dtype == np.float64 else 1e-18
        k = int(g[f"{name}/iterations"])
        qa = float(g[f"{name}/QA_cost"])
        mv = ex.matvec(kid, X, g[f"{name}/q"], c["v"], qa, 1.0 / c["cost"], **pr)
        out[f"{name}/matvec"] = mv
        out[f"{name}/ref_matvec_err"] = np.asarray(np.max(np.abs(g[f"{name}/matvec_p"].astype(np.float64) - mv)) / np.max(np.abs(mv)))
        ek = ex.solve(kid, X, y, cost=c["cost"], eps=tiny, max_iter=k, **pr)
        es = ex.solve(kid, X, y, cost=c["cost"], eps=1e-16 if X.dtype == np.float64 else 1e-16, max_iter=20 * X.shape[0], **pr)
        out[f"{name}/iterations"] = np.asarray(k)
        out[f"{name}/alpha_k"], out[f"{name}/rho_k"] = ek["alpha"], np.asarray(ek["rho"])
        out[f"{name}/alpha_star"], out[f"{name}/rho_star"] = es["alpha"], np.asarray(es["rho"])
        out[f"{name}/star_iterations"] = np.asarray(es["iterations"])
        errs = {"alpha_k": 0.0, "rho_k": 0.0, "alpha_star": 0.0, "rho_star": 0.0}

        def account(alpha, rho):
            for tag, e in (("k", ek), ("star", es)):
                sc = float(np.max(np.abs(e["alpha"])))
                errs[f"alpha_{tag}"] = max(errs[f"alpha_{tag}"], float(np.max(np.abs(alpha[:-1].astype(np.float64) - e["alpha"][:-1]))) / sc)
                errs[f"rho_{tag}"] = max(errs[f"rho_{tag}"], abs(float(rho) - e["rho"]))

        account(g[f"{name}/alpha"], g[f"{name}/rho"])
        for nthr in (8, 1, 3, 5):
            ref.set_threads(nthr)
            r = ref.solve(kid, X, y, cost=c["cost"], eps=c["eps"], max_iter=c["max_iter"], **pr)
            if r["iterations"] == k:
                account(r["alpha"], r["rho"])
        ref.set_threads(8)
        for key, val in errs.items():
            out[f"{name}/ref_{key.split('_')[0]}_err_{key.split('_')[1]}"] = np.asarray(val)
        pv = ex.predict(kid, X, g[f"{name}/alpha"], float(g[f"{name}/rho"]), c["P"], **pr)
        out[f"{name}/predict"] = pv
        out[f"{name}/ref_predict_err"] = np.asarray(np.max(np.abs(g[f"{name}/predict"].astype(np.float64) - pv)) / np.max(np.abs(pv)))
        print(f"{name:28s} k {k:3d} (x* after {es['iterations']}): reference errors  matvec {float(out[f'{name}/ref_matvec_err']):.1e}  alpha_k {errs['alpha_k']:.1e}  rho_k {errs['rho_k']:.1e}  "
              f"alpha* {errs['alpha_star']:.1e}  rho* {errs['rho_star']:.1e}  predict {float(out[f'{name}/ref_predict_err']):.1e}")
    np.savez_compressed(os.path.join(HERE, "exact_vectors.npz"), **out)


if __name__ == "__main__":
    main()
