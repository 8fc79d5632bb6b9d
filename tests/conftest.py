import os
import sys

import pytest

# tests/test_gpu_group.py puts two shards of one matrix on ONE GPU (their kernels wait for each other): they need
# their own hardware work queues, and no kernel may be loaded lazily (a load can synchronise the device) while
# the other shard is already waiting.  Must be in the environment before the first CUDA call of the process.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def hb():
    """The product package (hssmatrices.jl_b200 via the hssb200 shim), with the library built."""
    import hssb200
    if not os.path.exists(hssb200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    hssb200.lib()
    return hssb200


@pytest.fixture(scope="session")
def oracle():
    import hss_oracle
    return hss_oracle


@pytest.fixture(scope="session")
def ulv_oracle():
    import hss_ulv_oracle
    return hss_ulv_oracle
