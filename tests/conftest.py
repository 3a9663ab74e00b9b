import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# TEST INFRASTRUCTURE: tests/test_phys_cpu.py re-runs the column-physics GPU tests in a subprocess against the HOST build of the same
# sources (tests/host/build_phys_cpu.py).  That subprocess sets this variable to the path of libisca_phys_cpu.so; the ctypes mirrors
# of the test process are then bound to it and the gpu-marked tests are not skipped.  The isca_b200 package itself never looks at it.
CPU_BUILD = os.environ.get("ISCA_B200_TESTS_ON_CPU_BUILD")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    if CPU_BUILD:
        import ctypes
        from isca_b200 import api
        api._lib = ctypes.CDLL(CPU_BUILD)


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) the in-tree shared library; nvcc cross-compiles without a GPU."""
    from isca_b200 import build
    return build.build()


def _cuda_device_present() -> bool:
    if CPU_BUILD:
        return True
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a kernel that never returns) must end the run instead of holding the box until the outer limit:
    every gpu-marked test gets a pytest-timeout limit (thread method: works while the main thread sits inside a CUDA call)."""
    have_gpu = _cuda_device_present()
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests run on the B200 box (`pytest -m gpu`)")
    for item in items:
        if item.get_closest_marker("gpu") is not None and not have_gpu:
            item.add_marker(skip)
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(900, method="thread"))
