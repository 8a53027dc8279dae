import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.build()
    return o


@pytest.fixture(scope="session")
def gpu_ctx():
    """A solver context on cuda:0 through the C-ABI. No fallback: fails when the library or the
    device is missing."""
    from rspl_slam_b200 import build, capi
    build.build_library()
    ctx = capi.Context(device=0)
    yield ctx
    ctx.close()
