"""The reference's OWN command line front-ends (src/main_train.cpp, src/main_predict.cpp, compiled unmodified by
integration/Makefile with the b200 backend registered as `backend_type::b200`): `plssvm-train -b b200` / `plssvm-predict -b b200`
must behave like `-b openmp` — same LIBSVM model format, interchangeable models, same predictions."""
import os
import subprocess
import sys

import numpy as np
import pytest

from datagen import make_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_ref")


def _need_cli():
    if not os.path.exists(os.path.join(BIN, "plssvm-train")):
        pytest.skip("integration/_ref/plssvm-train not built (needs /root/reference)")


def _write_libsvm(path, X, y):
    with open(path, "w") as f:
        for xi, yi in zip(X, y):
            f.write(f"{int(yi)} " + " ".join(f"{j + 1}:{v:.17g}" for j, v in enumerate(xi)) + "\n")


def _run(*args):
    return subprocess.run(list(args), capture_output=True, text=True, timeout=300)


def test_cli_lists_and_parses_the_b200_backend(tmp_path):
    _need_cli()
    res = _run(os.path.join(BIN, "plssvm-train"), "--help")
    assert "openmp|b200" in res.stdout
    X, y = make_data(120, 8, 801)
    _write_libsvm(tmp_path / "train.libsvm", X, y)
    ok = _run(os.path.join(BIN, "plssvm-train"), "-b", "openmp", "-t", "2", "-e", "1e-8", "-q", str(tmp_path / "train.libsvm"), str(tmp_path / "ref.model"))
    assert ok.returncode == 0, ok.stderr
    assert "kernel_type rbf" in open(tmp_path / "ref.model").read()
    import torch
    if not torch.cuda.is_available():  # no CPU fallback: the backend must fail loudly, through the reference's own error path
        bad = _run(os.path.join(BIN, "plssvm-train"), "-b", "b200", "-t", "2", "-q", str(tmp_path / "train.libsvm"), str(tmp_path / "b200.model"))
        assert bad.returncode != 0 and "b200::backend_exception" in bad.stderr
    unknown = _run(os.path.join(BIN, "plssvm-train"), "-b", "nonsense", str(tmp_path / "train.libsvm"))
    assert unknown.returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["0", "1", "2"])
def test_reference_cli_with_the_b200_backend(kernel, tmp_path):
    _need_cli()
    X, y = make_data(800, 40, 810 + int(kernel))
    P, yP = make_data(300, 40, 820 + int(kernel))
    _write_libsvm(tmp_path / "train.libsvm", X, y)
    _write_libsvm(tmp_path / "test.libsvm", P, yP)
    models = {}
    for backend in ("openmp", "b200"):
        models[backend] = str(tmp_path / f"{backend}.model")
        res = _run(os.path.join(BIN, "plssvm-train"), "-b", backend, "-t", kernel, "-e", "1e-8", "-q", str(tmp_path / "train.libsvm"), models[backend])
        assert res.returncode == 0, res.stderr
    header = {b: [line for line in open(m).read().split("SV\n")[0].splitlines() if not line.startswith("#") and not line.startswith("rho")] for b, m in models.items()}
    assert header["openmp"] == header["b200"]  # identical LIBSVM model header (kernel, gamma, labels, nr_sv) up to rho's last digits
    preds = {}
    for model_backend, model in models.items():
        for backend in ("openmp", "b200"):
            out = str(tmp_path / f"{model_backend}_{backend}.predict")
            res = _run(os.path.join(BIN, "plssvm-predict"), "-b", backend, str(tmp_path / "test.libsvm"), model, out)
            assert res.returncode == 0, res.stderr
            assert "Accuracy" in res.stdout
            preds[(model_backend, backend)] = np.loadtxt(out).astype(int)
    # Same model file, either backend -> IDENTICAL labels except where the decision value is at rounding level.  Decision values: the model file
    # (alpha with 11 significant digits, class-grouped support vectors) evaluated in extended precision (oracle/lssvm_exact.cpp).
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import oracle
    from make_golden import parse_model
    from parity import assert_same_labels_outside_band
    ex = oracle.Exact()
    f = {}
    for model_backend, model in models.items():
        m = parse_model(model)
        f[model_backend] = ex.predict(m["kernel"], m["sv"], m["alpha"], m["rho"], P, m["degree"], m["gamma"], m["coef0"])
        band = 1e-9 * float(np.sum(np.abs(m["alpha"])))
        assert band < 1e-4 * float(np.max(np.abs(f[model_backend])))  # the band is a sliver of the value range: labels are identical in practice
        assert_same_labels_outside_band(preds[(model_backend, "openmp")], preds[(model_backend, "b200")], f[model_backend], band, f"kernel {kernel}, {model_backend} model")
    # different training backend: the models differ by the CG noise of either solve; labels identical outside 10 x the deviation of their decision values
    band = 10.0 * float(np.max(np.abs(f["openmp"] - f["b200"])))
    assert_same_labels_outside_band(preds[("openmp", "openmp")], preds[("b200", "b200")], f["openmp"], band, f"kernel {kernel}, cross")


def _cg_section(path):
    """The `cg:` block of the reference's performance-tracker YAML (src/plssvm/detail/performance_tracker.cpp) as a dict of strings."""
    out, inside = {}, False
    for line in open(path).read().splitlines():
        if line.startswith("cg:"):
            inside = True
        elif inside and line.startswith("  "):
            k, v = line.strip().split(":", 1)
            out[k.strip()] = v.strip()
        elif inside:
            break
    return out


@pytest.mark.gpu
def test_logging_and_tracker_contract_like_the_reference(tmp_path):
    """SURVEY.md §8(b) logging contract (gpu_csvm.hpp:569-571, 637-646): `plssvm-train -b b200` prints the reference's per-iteration, summary and
    LIBSVM-style lines and fills the same `cg.*` performance-tracker entries as `-b openmp`."""
    import re
    _need_cli()
    X, y = make_data(600, 24, 830)
    _write_libsvm(tmp_path / "train.libsvm", X, y)
    out, cg = {}, {}
    for backend in ("openmp", "b200"):
        track = str(tmp_path / f"{backend}.yaml")
        res = _run(os.path.join(BIN, "plssvm-train"), "-b", backend, "-t", "2", "-e", "1e-8", "--performance_tracking", track, str(tmp_path / "train.libsvm"),
                   str(tmp_path / f"{backend}.model"))
        assert res.returncode == 0, res.stderr
        out[backend], cg[backend] = res.stdout, _cg_section(track)
    assert set(cg["b200"]) >= {"iterations", "max_iterations", "residuum", "target_residuum", "avg_iteration_time", "epsilon"}
    assert set(cg["b200"]) >= set(cg["openmp"]) - {"total_runtime"} or set(cg["b200"]) >= set(cg["openmp"])
    it = {b: int(cg[b]["iterations"]) for b in cg}
    assert abs(it["b200"] - it["openmp"]) <= 1 and cg["b200"]["max_iterations"] == cg["openmp"]["max_iterations"] == "600"
    assert float(cg["b200"]["epsilon"]) == float(cg["openmp"]["epsilon"]) == 1e-8
    assert abs(float(cg["b200"]["target_residuum"]) / float(cg["openmp"]["target_residuum"]) - 1.0) < 1e-9
    assert float(cg["b200"]["residuum"]) <= float(cg["b200"]["target_residuum"])
    assert cg["b200"]["avg_iteration_time"].endswith("ms")
    line = re.compile(r"Start Iteration (\d+) \(max: 600\) with current residuum (\S+) \(target: (\S+)\)\. Done in \d+ms\.")
    got = {b: line.findall(out[b]) for b in out}
    assert len(got["b200"]) == it["b200"] and [int(g[0]) for g in got["b200"]] == list(range(1, it["b200"] + 1))
    # the residual history follows the reference's while rounding noise is not yet amplified
    for (_, r_b, t_b), (_, r_o, t_o) in list(zip(got["b200"], got["openmp"]))[:3]:
        assert abs(float(r_b) / float(r_o) - 1.0) < 1e-8 and abs(float(t_b) / float(t_o) - 1.0) < 1e-9
    assert re.search(rf"Finished after {it['b200']}/600 iterations with a residuum of \S+ \(target: \S+\) and an average iteration time of \d+ms\.", out["b200"])
    assert "Using B200 as backend." in out["b200"] and re.search(r"Found \d+ B200 device\(s\)", out["b200"])
    res = _run(os.path.join(BIN, "plssvm-train"), "-b", "b200", "-t", "2", "-e", "1e-8", "--verbosity", "libsvm", str(tmp_path / "train.libsvm"), str(tmp_path / "l.model"))
    assert res.returncode == 0 and f"optimization finished, #iter = {it['b200']}" in res.stdout
