import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """The reference-compiled oracle build; prebuilt in the build container and shipped in oracle/_ref/."""
    import oracle
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
    return oracle.Oracle("reference")
