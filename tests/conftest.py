import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pycuda-euler_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import _native
    c = _native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def g200_reads():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "g200.json")) as f:
        return json.load(f)["reads"]
