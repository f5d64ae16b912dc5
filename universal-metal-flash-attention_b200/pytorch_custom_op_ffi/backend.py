"""Backend registration and context management -- mirror of the reference's
examples/pytorch-custom-op-ffi/python/pytorch_custom_op_ffi/backend.py (same functions, same error behaviour).

The reference renames PyTorch's PrivateUse1 device and registers a dispatcher override in C++
(src/metal_sdpa_backend.cpp:1906-1990).  CUDA tensors already have a device, so registering here installs a routing
wrapper around torch.nn.functional.scaled_dot_product_attention: calls whose tensors live on the library's CUDA device
with a supported dtype / rank go to libMFAFFI.so (ext.metal_scaled_dot_product_attention); everything else reaches
PyTorch's own implementation untouched and is counted in get_dispatch_stats()["fallback_native"].
"""
import threading
import warnings
from contextlib import contextmanager
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import ext as _ext

_backend_registered = False
_registration_lock = threading.Lock()
_native_sdpa = F.scaled_dot_product_attention


def is_metal_sdpa_available() -> bool:
    """True when libMFAFFI.so sees a supported (sm_100) device."""
    try:
        return torch.cuda.is_available() and _ext.is_metal_available()
    except Exception:
        return False


def metal_sdpa_version() -> Optional[Tuple[int, int, int]]:
    if not is_metal_sdpa_available():
        return None
    return _ext.get_version()


def _routable(query, key, value, dropout_p) -> bool:
    if not (isinstance(query, torch.Tensor) and query.is_cuda and key.is_cuda and value.is_cuda):
        return False
    if query.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        return False
    if key.device != query.device or value.device != query.device:
        return False                     # operands spread over devices: PyTorch's own path raises the proper error
    if not _ext.device_supported(query.device):
        return False                     # a GPU in this process the library has no kernels for (not sm_100)
    if key.dtype != query.dtype or value.dtype != query.dtype:
        return False
    if query.dim() < 2 or query.dim() > 4 or key.dim() != query.dim() or value.dim() != query.dim():
        return False
    if query.size(-1) > 256 or query.size(-1) != key.size(-1) or value.size(-1) != query.size(-1):
        return False
    if dropout_p and dropout_p > 0.0:
        return False
    if query.dim() >= 3 and query.size(-3) != key.size(-3) and (key.size(-3) == 0 or query.size(-3) % key.size(-3)):
        return False
    return True


def _routed_sdpa(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False):
    if _routable(query, key, value, dropout_p):
        return _ext.metal_scaled_dot_product_attention(query, key, value, attn_mask, dropout_p, is_causal, scale,
                                                       enable_gqa)
    _ext._stats["total"] += 1
    _ext._stats["fallback_native"] += 1
    return _native_sdpa(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
                        enable_gqa=enable_gqa)


def register_metal_sdpa_backend() -> None:
    """Route torch.nn.functional.scaled_dot_product_attention through the B200 engine."""
    global _backend_registered
    with _registration_lock:
        if _backend_registered:
            warnings.warn("Metal SDPA backend already registered", UserWarning)
            return
        if not is_metal_sdpa_available():
            raise RuntimeError("Metal is not available on this device")
        F.scaled_dot_product_attention = _routed_sdpa
        torch.nn.functional.scaled_dot_product_attention = _routed_sdpa
        _backend_registered = True


def unregister_metal_sdpa_backend() -> None:
    global _backend_registered
    with _registration_lock:
        if not _backend_registered:
            warnings.warn("Metal SDPA backend not currently registered", UserWarning)
            return
        F.scaled_dot_product_attention = _native_sdpa
        torch.nn.functional.scaled_dot_product_attention = _native_sdpa
        _backend_registered = False


@contextmanager
def use_metal_sdpa():
    """Temporarily enable the backend; yields the execution device."""
    was_registered = _backend_registered
    if not was_registered:
        register_metal_sdpa_backend()
    try:
        yield _resolve_execution_device()
    finally:
        if not was_registered:
            unregister_metal_sdpa_backend()


class MetalSDPAContext:
    """Fine-grained control (reference backend.py:140-230): to_device / to_cpu / direct_call."""

    def __init__(self, auto_register: bool = True):
        self.auto_register = auto_register
        self.device = None
        self._registered_here = False

    def __enter__(self):
        if self.auto_register and not _backend_registered:
            register_metal_sdpa_backend()
            self._registered_here = True
        self.device = _resolve_execution_device()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.device = None
        if self._registered_here:
            unregister_metal_sdpa_backend()
            self._registered_here = False

    def to_device(self, tensor: torch.Tensor) -> torch.Tensor:
        if self.device is None:
            raise RuntimeError("Context not active")
        return tensor.to(self.device)

    def to_cpu(self, tensor: torch.Tensor) -> torch.Tensor:
        return tensor.cpu()

    def direct_call(self, query, key, value, attn_mask=None, dropout_p: float = 0.0, is_causal: bool = False,
                    scale: Optional[float] = None) -> torch.Tensor:
        """Bypass the routing wrapper and call the engine; the result comes back on the caller's device / dtype.
        (The reference promotes fp16 / bf16 to fp32 torch SDPA here, backend.py:202-213; the B200 kernels take them
        natively.)"""
        if self.device is None:
            raise RuntimeError("MetalSDPAContext is not active")
        orig_device, orig_dtype = query.device, query.dtype
        q, k, v = (t.to(self.device) for t in (query, key, value))
        m = attn_mask.to(self.device) if attn_mask is not None else None
        result = _ext.metal_scaled_dot_product_attention(q, k, v, m, dropout_p, is_causal, scale)
        if result.device != orig_device:
            result = result.to(orig_device)
        if result.dtype != orig_dtype:
            result = result.to(orig_dtype)
        return result


class MetalSDPABackendConfig:
    """torch.backends.metal_sdpa (enabled / available / version), as in the reference."""

    @property
    def enabled(self) -> bool:
        return _backend_registered

    @enabled.setter
    def enabled(self, value: bool):
        if value and not _backend_registered:
            register_metal_sdpa_backend()
        elif not value and _backend_registered:
            unregister_metal_sdpa_backend()

    @property
    def available(self) -> bool:
        return is_metal_sdpa_available()

    @property
    def version(self) -> Optional[Tuple[int, int, int]]:
        return metal_sdpa_version()


if not hasattr(torch.backends, "metal_sdpa"):
    torch.backends.metal_sdpa = MetalSDPABackendConfig()


def _resolve_execution_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    raise RuntimeError("Metal SDPA backend requires an available CUDA (sm_100) device")
