import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# GPU run order: the hot path first (kernels, solves, the configs of BASELINE.json against the reference library), then the chains
# either side of it, then the keyword / sos_proc front ends.  With -x a failure at the edge does not hide the parity of the core.
_GPU_ORDER = ("test_gpu_parity.py", "test_gpu_vs_reference.py", "test_profile_chain.py", "test_aerosol_chain.py", "test_writers.py",
              "test_frontend.py", "test_surface_nadal.py")


def pytest_collection_modifyitems(config, items):
    def key(it):
        if it.get_closest_marker("gpu") is None:
            return (0, 0)
        name = os.path.basename(str(it.fspath))
        return (1, _GPU_ORDER.index(name) if name in _GPU_ORDER else len(_GPU_ORDER))
    items.sort(key=key)                                  # stable: order within a file is kept


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("radiativetransfer-sos_b200")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def solver(pkg):
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    s = api.Solver(0)          # raises when the CUDA library / device is missing: no fallback
    yield s
    s.close()
