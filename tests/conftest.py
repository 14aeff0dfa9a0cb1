import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_binding as ob
    return ob.Oracle()


@pytest.fixture(scope="session")
def reference():
    from tests import oracle_binding as ob
    if not ob.Reference.available:
        pytest.skip("oracle/_ref/libvgref.so not built (no /root/reference on this machine)")
    return ob.Reference()


@pytest.fixture(scope="session")
def vglib():
    """The product library; built on demand when the .so is stale or missing."""
    from varigraph_b200 import build
    build.build()
    from varigraph_b200 import capi
    return capi


@pytest.fixture(scope="session")
def ctx(vglib):
    c = vglib.Context(0, buffer_mb=1)  # 1 MiB staging chunks: even small tests cross chunk boundaries
    yield c
    c.close()
