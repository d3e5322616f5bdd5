import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """CUDA context through the C ABI.  No skip-on-missing: a GPU test without a GPU or without
    the built library must fail loudly (there is no CPU fallback to fall back to)."""
    from xivo_b200 import Context

    c = Context(0)
    yield c
    c.close()
