import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("urlearning-cpp_b200")


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def engine(pkg):
    eng = pkg.Engine(0)  # raises loudly when the library or the device is missing
    yield eng
    eng.close()


@pytest.fixture(scope="session", params=["cube", "tree", "direct"])
def bic_engine(pkg, request):
    """BIC engines for the K1 strategies: 'cube' (the default: roots counted into global tables, the rest marginalised through
    HBM), 'tree' (tables counted in shared-memory slices of bucketed packed rows, subtrees marginalised on chip; falls back
    to cube when a family cannot be laid out that way) and 'direct' (every set counted from the rows)."""
    old = os.environ.get("URLGPU_BIC_MODE")
    os.environ["URLGPU_BIC_MODE"] = request.param
    try:
        eng = pkg.Engine(0)
    finally:
        if old is None:
            os.environ.pop("URLGPU_BIC_MODE", None)
        else:
            os.environ["URLGPU_BIC_MODE"] = old
    yield eng
    eng.close()


DATA = os.path.join(ROOT, "tests", "data")


@pytest.fixture(scope="session")
def data_dir():
    return DATA
