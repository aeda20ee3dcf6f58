import os
import sys

import pytest

# the single-process sharded tests step 8 handles from 8 threads: give their streams separate hardware queues
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_lib():
    """The CPU restatement (oracle/wembed_port.cpp), compiled on demand."""
    import oracle
    oracle.build("port")
    return "port"


@pytest.fixture(scope="session")
def ref_lib():
    """The reference's own sources (oracle/_ref); skipped where neither the checkout nor a prebuilt .so exists."""
    import oracle
    if not oracle.have("ref") and not oracle.build("ref"):
        pytest.skip("reference checkout not available: oracle/_ref cannot be built here")
    elif os.path.isdir("/root/reference"):
        oracle.build("ref")
    return "ref"


@pytest.fixture(scope="session")
def device_lib():
    from wembed_b200 import build, cabi
    build.build()
    l = cabi.lib()
    if l.wb_device_count() < 1:
        pytest.skip("no CUDA device")
    return cabi
