import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the host libraries and the oracle if a fresh checkout has none (CPU-only, seconds).
    libptcuda.so is only (re)built when nvcc is present; GPU tests fail loudly without it."""
    from pathtracer_ocl_b200 import build as b
    b.build_scene_lib()
    try:
        b.build_cuda()
    except RuntimeError:
        pass
    from oracle import oracle as O
    O.build()
    yield


def has_gpu() -> bool:
    try:
        from pathtracer_ocl_b200 import trace as T
        return T.lib().ptc_device_count() > 0
    except Exception:
        return False
