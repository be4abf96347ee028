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


def engine_with_env(pkg, env):
    """an Engine created under the given environment overrides (the library reads its strategy switches at urlgpu_create)"""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return pkg.Engine(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(scope="session", params=["cube", "tree", "direct", "cube-rank", "tree-rank", "direct-rank"])
def bic_engine(pkg, request):
    """BIC engines for the K1 strategies: 'cube' (the default: roots counted into global tables, the rest marginalised through
    HBM), 'tree' (tables counted in shared-memory slices of bucketed packed rows, subtrees marginalised on chip; falls back
    to cube when a family cannot be laid out that way) and 'direct' (every set counted from the rows); '-rank': the same
    strategies writing into the rank-space (colex) layout of the score cache instead of the dense 2^c table."""
    mode, _, layout = request.param.partition("-")
    eng = engine_with_env(pkg, {"URLGPU_BIC_MODE": mode, "URLGPU_LAYOUT": layout or "dense"})
    yield eng
    eng.close()


@pytest.fixture(scope="session", params=["dense", "rank"])
def cbic_engine(pkg, request):
    """cBIC engines for the two cache layouts: 'dense' (2^c table by compact mask: level-A sweeps + DFS, segment DPs) and
    'rank' (colex-rank table: per-set sweeps, per-layer DPs; used whenever the parent limit allows, else dense)."""
    eng = engine_with_env(pkg, {"URLGPU_LAYOUT": request.param})
    yield eng
    eng.close()


DATA = os.path.join(ROOT, "tests", "data")


@pytest.fixture(scope="session")
def data_dir():
    return DATA
