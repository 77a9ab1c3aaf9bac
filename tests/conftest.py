import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import gsc_oracle
    gsc_oracle.build()
    return gsc_oracle


@pytest.fixture(scope="session")
def lib_built():
    from soundchunks_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def ctx(lib_built):
    import soundchunks_b200 as sc
    c = sc.Context(0)
    yield c
    c.close()
