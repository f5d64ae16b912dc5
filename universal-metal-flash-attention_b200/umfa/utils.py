"""Small helpers mirroring examples/python-ffi/src/umfa/utils.py of the reference."""
import ctypes
from typing import Tuple

from ._ffi import _lib
from .core import MFAContext


def is_metal_available() -> bool:
    """Name kept for drop-in compatibility: True iff the library can run (an sm_100 GPU is visible)."""
    try:
        return bool(_lib.mfa_is_device_supported())
    except Exception:
        return False


is_device_available = is_metal_available


def get_version() -> Tuple[int, int, int]:
    major, minor, patch = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.mfa_get_version(ctypes.byref(major), ctypes.byref(minor), ctypes.byref(patch))
    return major.value, minor.value, patch.value


def create_context() -> MFAContext:
    if not is_metal_available():
        raise RuntimeError("No supported GPU (sm_100) is available")
    return MFAContext()


def print_system_info():
    print("libMFAFFI (B200) version", ".".join(map(str, get_version())))
    print("device supported:", is_metal_available(), "| devices:", _lib.mfa_get_device_count())
