import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # parametrised engine ids ending in "cuda" are GPU tests
    for item in items:
        if "engine" in getattr(item, "fixturenames", ()) and "[cuda" in item.nodeid or "-cuda" in item.nodeid:
            item.add_marker(pytest.mark.gpu)


def _engines():
    from tests import engines
    return [pytest.param(engines.OracleEngine(), id="oracle"),
            pytest.param(engines.CudaEngineLazy(), id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=_engines())
def engine(request):
    e = request.param
    e.ensure()
    return e
