import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_device_present():
    """A CUDA device, asked of the driver directly (no torch import: it costs ~a minute on a fresh box)."""
    import ctypes

    for name in ("libcuda.so.1", "libcuda.so"):
        try:
            cu = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a GPU SKIPS the gpu-marked tests instead of failing in them (the same tests
    run on the host warp emulator through tests/test_kernels_emulated.py, where they carry no gpu mark)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: libgwi has no CPU fallback (the emulated copies of these tests run instead)")
    for it in gpu_items:
        it.add_marker(skip)
