import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """Test-only CPU emulator of the kernel bodies (same sources, -DSPIM_HOST_EMU)."""
    import __graft_entry__ as g
    from spim_registration_b200 import native
    return native.load_library(g.build_emulator())


@pytest.fixture(scope="session")
def cuda_lib():
    from spim_registration_b200 import build, native
    return native.load_library(build.build_cuda_library())


@pytest.fixture(scope="session")
def gpu(cuda_lib):
    n = cuda_lib.getNumDevicesCUDA()
    if n <= 0:
        pytest.fail(f"no CUDA device visible (getNumDevicesCUDA() = {n}); GPU tests have no CPU fallback")
    return cuda_lib
