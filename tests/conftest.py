import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# tests/test_gpu_group.py runs up to 8 partitioned solvers of this process on one GPU, each with its own streams and with kernels
# that wait for flags raised by another solver's kernels: give every stream its own hardware queue (read at CUDA initialisation)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def afx():
    import aeroflex_b200
    aeroflex_b200.load_library()
    return aeroflex_b200


@pytest.fixture(scope="session")
def gpu(afx):
    n = afx.device_count()
    if n <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible (the product has no CPU fallback)")
    return n
