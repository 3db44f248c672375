import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """the CUDA library and the oracle are built in-tree; build them when missing (CPU box) -- never fall back"""
    from fraytracer_b200 import _abi
    from oracle import oracle as O

    if not os.path.exists(_abi.LIB_PATH) or not os.path.exists(O.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    yield


def bits(a):
    import numpy as np

    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
