import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle (g++) and the product library (nvcc, cross-compiles without a GPU) are built once per session."""
    from oracle import oracle as O
    from raymond_b200 import build as B
    O.build()
    B.build()
    yield
