"""The reference's OWN command line front-ends (src/main_train.cpp, src/main_predict.cpp, compiled unmodified by
integration/Makefile with the b200 backend registered as `backend_type::b200`): `plssvm-train -b b200` / `plssvm-predict -b b200`
must behave like `-b openmp` — same LIBSVM model format, interchangeable models, same predictions."""
import os
import subprocess

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
    for model_backend in models:  # same model file, either backend -> same labels
        assert (preds[(model_backend, "openmp")] != preds[(model_backend, "b200")]).mean() <= 0.005
    assert (preds[("openmp", "openmp")] != preds[("b200", "b200")]).mean() <= 0.01  # different training backend: CG noise band only
