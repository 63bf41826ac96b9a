"""The C++ host adaptor (include/plssvm_b200/csvm.hpp) compiles against the C ABI with a plain g++ and behaves like a
reference backend: without a GPU the constructor throws backend_exception, on the B200 the reference's trivial tests pass."""
import os
import subprocess

import pytest

import plssvm_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_adaptor")


def _build():
    pb.load_library()
    libdir = os.path.dirname(pb.lib_path())
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_adaptor.cpp"),
           "-o", BIN, "-L", libdir, "-lplssvm_b200", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return BIN


def test_adaptor_compiles_and_fails_loudly_without_a_device():
    import torch
    exe = _build()
    res = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert res.returncode == 0, res.stdout + res.stderr
    else:
        assert res.returncode == 3, res.stdout + res.stderr
        assert "backend_exception" in res.stdout


@pytest.mark.gpu
def test_adaptor_passes_the_reference_trivial_tests():
    exe = _build()
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "adaptor tests passed" in res.stdout
