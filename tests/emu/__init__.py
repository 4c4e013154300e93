"""TEST INFRASTRUCTURE: host warp emulator of the CUDA kernels (see tests/emu/cuda_runtime.h).

``activate()`` points the ctypes binding (gwinferno_b200.capi) at tests/emu/libgwi_emu.so -- the
kernel sources of gwinferno_b200/csrc compiled by g++ with every CUDA thread run as a fiber -- and
``deactivate()`` restores the product library.  Only tests call this; the product never does and
``capi.load_library()`` refuses an emulator build.
"""

import ctypes as C
import os
import subprocess

from gwinferno_b200 import capi
from gwinferno_b200 import likelihood as _likelihood

HERE = os.path.dirname(os.path.abspath(__file__))
_saved = None


def library_path(variant=""):
    return os.path.join(HERE, f"libgwi_emu_{variant}.so" if variant else "libgwi_emu.so")


def build(variant="", extra=""):
    """g++ build of the emulator library (a few seconds per translation unit, parallel)."""
    cmd = ["make", "-C", HERE, f"-j{max(4, min(16, os.cpu_count() or 4))}"]
    if variant:
        cmd += [f"VARIANT={variant}", f"EXTRA={extra}"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return library_path(variant)


def activate(variant=""):
    """Route capi (library + device buffers) to the emulator build.  Returns the library."""
    global _saved
    path = library_path(variant)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: build it with tests.emu.build()")
    lib = C.CDLL(path)
    assert lib.gwi_emu_marker() == 1
    if _saved is None:
        _saved = (capi._lib, capi._cudart, capi.LIB_PATH)
    _likelihood.clear_cache()
    capi._lib = None
    capi._cudart = None
    capi.LIB_PATH = path
    capi.load_library(_allow_emulator=True)
    # the emulator library exports the handful of CUDA runtime calls the binding uses
    rt = lib
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaFree.argtypes = [C.c_void_p]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rt.cudaGetErrorString.restype = C.c_char_p
    rt.cudaGetErrorString.argtypes = [C.c_int]
    capi._cudart = rt
    return lib


def deactivate():
    global _saved
    if _saved is None:
        return
    _likelihood.clear_cache()
    capi._lib, capi._cudart, capi.LIB_PATH = _saved
    _saved = None
