"""CPU checks of the PyTorch adapter twin: public names of the reference package exist
(examples/pytorch-custom-op-ffi/python/pytorch_custom_op_ffi/__init__.py:22-39, python_bindings.cpp:40-170), and without
a CUDA device registration fails loudly instead of silently keeping PyTorch's path."""
import pytest


def test_public_names():
    import pytorch_custom_op_ffi as p
    from pytorch_custom_op_ffi import ext
    for name in ("register_metal_sdpa_backend", "unregister_metal_sdpa_backend", "use_metal_sdpa",
                 "is_metal_sdpa_available", "metal_sdpa_version", "MetalSDPAContext"):
        assert hasattr(p, name)
    for name in ("metal_scaled_dot_product_attention", "rope_scaled_dot_product_attention",
                 "metal_flash_attention_autograd", "metal_quantized_flash_attention_autograd", "set_quantization_mode",
                 "clear_quantization_mode", "get_dispatch_stats", "reset_dispatch_stats", "hadamard_rotate",
                 "quantized_scaled_dot_product_attention", "is_metal_available", "has_native_bfloat", "get_version"):
        assert callable(getattr(ext, name))
    assert (ext.QUANT_INT8, ext.QUANT_INT4, ext.QUANT_TENSOR_WISE, ext.QUANT_BLOCK_WISE) == (3, 4, 0, 2)
    assert ext.get_version() == (1, 0, 0)


def test_no_silent_fallback_without_device():
    import torch
    import pytorch_custom_op_ffi as p
    from pytorch_custom_op_ffi import ext
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    assert not p.is_metal_sdpa_available()
    with pytest.raises(RuntimeError):
        p.register_metal_sdpa_backend()
    q = torch.randn(1, 1, 8, 8)
    with pytest.raises(RuntimeError):
        ext.metal_scaled_dot_product_attention(q, q, q)          # CPU tensors are refused, never computed on the host
    with pytest.raises(RuntimeError):
        ext.set_quantization_mode(7, 0)
    assert hasattr(torch.backends, "metal_sdpa") and torch.backends.metal_sdpa.enabled is False
