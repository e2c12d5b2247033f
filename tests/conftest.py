import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def cuda_device_count() -> int:
    """CUDA devices visible to this process (0 on a CPU box), asked of the runtime the product links — no torch import."""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped, not failed, on a box without a CUDA device (plain `pytest` in the build container)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and cuda_device_count() == 0:
        skip = pytest.mark.skip(reason="no CUDA device: the product has no CPU path")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle (g++) and the product library (nvcc, cross-compiles without a GPU) are built once per session."""
    from oracle import oracle as O
    from raymond_b200 import build as B
    O.build()
    B.build()
    yield
