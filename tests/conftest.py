import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def restatement():
    from oracle import pyoracle
    return pyoracle.Restatement()


@pytest.fixture(scope="session")
def reference():
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_SO) and not os.path.exists("/root/reference/swgl.c"):
        pytest.skip("compiled reference (oracle/_ref) not available")
    return pyoracle.Reference()


@pytest.fixture(scope="session")
def gpu_api():
    import swgl_b200
    return swgl_b200.load()
