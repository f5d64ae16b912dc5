"""GPU tests of the CUDA twin of the reference's PyTorch adapter (pytorch_custom_op_ffi), modelled on
examples/pytorch-custom-op-ffi/tests/test_backend.py and tests/conftest.py:147-250 there: the routed
F.scaled_dot_product_attention is compared with PyTorch's own math SDPA in fp32 on the same inputs
(tolerances from the reference's conftest.py:189-198: fp32 rtol 1e-5 / atol 1e-6 is for its fp32 Metal path; here
fp32 1e-5 relative to max|ref|, fp16 1e-3, bf16 1e-2 on O; gradients 2e-2 for 16-bit inputs), and against the CPU
oracle for one causal case."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


@pytest.fixture(scope="module")
def adapter():
    import pytorch_custom_op_ffi as p
    from pytorch_custom_op_ffi import ext
    return p, ext


def ref_sdpa(torch, q, k, v, mask=None, causal=False, scale=None):
    """fp32 math reference written out (no fused kernels)."""
    q32, k32, v32 = q.float(), k.float(), v.float()
    if k32.size(1) != q32.size(1):
        g = q32.size(1) // k32.size(1)
        k32, v32 = k32.repeat_interleave(g, 1), v32.repeat_interleave(g, 1)
    s = (q32 @ k32.transpose(-1, -2)) * (scale if scale is not None else 1.0 / math.sqrt(q.size(-1)))
    if causal:
        Sq, Skv = s.shape[-2:]
        s = s.masked_fill(~torch.ones(Sq, Skv, dtype=torch.bool, device=q.device).tril(), float("-inf"))
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf")) if mask.dtype == torch.bool else s + mask.float()
    return torch.softmax(s, dim=-1) @ v32


def relmax(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30))


TOL = {"float32": 1e-5, "float16": 1e-3, "bfloat16": 1e-2}


@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("causal", [False, True])
def test_routed_sdpa_matches_math(torch_mod, adapter, dtype, causal):
    torch = torch_mod
    p, ext = adapter
    dt = getattr(torch, dtype)
    g = torch.Generator(device="cuda").manual_seed(7)
    q, k, v = (torch.randn(2, 4, 320, 64, device="cuda", generator=g).to(dt) for _ in range(3))
    ext.reset_dispatch_stats()
    with p.use_metal_sdpa():
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)
    assert out.dtype == dt and out.shape == q.shape
    assert relmax(out, ref_sdpa(torch, q, k, v, causal=causal)) < TOL[dtype]
    st = ext.get_dispatch_stats()
    assert st["total"] == 1 and st["direct"] == 1 and st["fallback_native"] == 0
    assert torch.nn.functional.scaled_dot_product_attention is not None and not torch.backends.metal_sdpa.enabled


def test_flux_shape_bf16_uses_tensor_core_kernel(torch_mod, adapter):
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(1, 24, 4608, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    with p.use_metal_sdpa():
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    assert ext._context(q.device).last_kernel == "fwd_tc_bf16_d128"
    ref = ref_sdpa(torch, q[:, :2], k[:, :2], v[:, :2])
    assert relmax(out[:, :2], ref) < 1e-2


@pytest.mark.parametrize("kind", ["bool", "additive", "broadcast", "all_true"])
def test_masks(torch_mod, adapter, kind):
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, Sq, Skv, D = 2, 3, 96, 160, 64
    q = torch.randn(B, H, Sq, D, device="cuda", generator=g)
    k, v = (torch.randn(B, H, Skv, D, device="cuda", generator=g) for _ in range(2))
    if kind == "bool":
        mask = torch.rand(B, H, Sq, Skv, device="cuda", generator=g) > 0.3
        mask[..., 0] = True
    elif kind == "additive":
        mask = torch.randn(B, H, Sq, Skv, device="cuda", generator=g)
    elif kind == "broadcast":
        mask = torch.randn(1, 1, Sq, Skv, device="cuda", generator=g)
    else:
        mask = torch.ones(B, 1, Sq, Skv, device="cuda", dtype=torch.bool)
    ext.reset_dispatch_stats()
    with p.use_metal_sdpa():
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    assert relmax(out, ref_sdpa(torch, q, k, v, mask=mask)) < 1e-5
    assert ext.get_dispatch_stats()["mask_all_true_skipped"] == (1 if kind == "all_true" else 0)


def test_low_rank_inputs_and_gqa(torch_mod, adapter):
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(11)
    q2, k2, v2 = (torch.randn(48, 32, device="cuda", generator=g) for _ in range(3))
    with p.use_metal_sdpa():
        o2 = torch.nn.functional.scaled_dot_product_attention(q2, k2, v2)
    assert o2.shape == (48, 32)
    assert relmax(o2, ref_sdpa(torch, q2[None, None], k2[None, None], v2[None, None])[0, 0]) < 1e-5
    q = torch.randn(1, 8, 64, 64, device="cuda", generator=g)
    k, v = (torch.randn(1, 2, 64, 64, device="cuda", generator=g) for _ in range(2))
    with p.use_metal_sdpa():
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, enable_gqa=True)
    assert relmax(o, ref_sdpa(torch, q, k, v)) < 1e-5


@pytest.mark.parametrize("dtype,causal", [("float32", False), ("float32", True), ("bfloat16", False), ("bfloat16", True)])
def test_autograd_matches_math(torch_mod, adapter, dtype, causal):
    torch = torch_mod
    p, ext = adapter
    dt = getattr(torch, dtype)
    g = torch.Generator(device="cuda").manual_seed(5)
    D = 128 if dtype == "bfloat16" else 64
    q, k, v = (torch.randn(1, 2, 256, D, device="cuda", generator=g).to(dt).requires_grad_(True) for _ in range(3))
    go = torch.randn(1, 2, 256, D, device="cuda", generator=g).to(dt)
    ext.reset_dispatch_stats()
    with p.use_metal_sdpa():
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)
    out.backward(go)
    assert ext.get_dispatch_stats()["fp32_autograd"] == 1
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ref = ref_sdpa(torch, qr, kr, vr, causal=causal)
    ref.backward(go.float())
    tol = 1e-4 if dtype == "float32" else 2e-2
    assert relmax(out, ref) < (1e-5 if dtype == "float32" else 1e-2)
    for a, b in ((q.grad, qr.grad), (k.grad, kr.grad), (v.grad, vr.grad)):
        assert a.dtype == dt
        assert relmax(a, b) < tol


@pytest.mark.parametrize("mask_kind", ["bf16_additive", "bool"])
def test_autograd_with_dense_mask_takes_the_staged_kernels(torch_mod, adapter, mask_kind):
    """F.scaled_dot_product_attention with a dense bf16 bias / bool mask in the tensors' own dtype: the mask is handed over in
    place and both directions run on the kernels that stage mask tiles in shared memory by TMA; O and the gradients against
    the written-out fp32 math"""
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(21)
    B, H, Sq, Skv, D = 1, 3, 384, 512, 128
    q = torch.randn(B, H, Sq, D, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
    k, v = (torch.randn(B, H, Skv, D, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True) for _ in range(2))
    go = torch.randn(B, H, Sq, D, device="cuda", generator=g).to(torch.bfloat16)
    if mask_kind == "bool":
        mask = torch.rand(1, H, Sq, Skv, device="cuda", generator=g) > 0.4
        mask[..., 0] = True
    else:
        mask = (1.5 * torch.randn(1, H, Sq, Skv, device="cuda", generator=g)).to(torch.bfloat16)
    with p.use_metal_sdpa():
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    assert ext._context(q.device).last_kernel == "fwd_tc_bf16_d128_tma_mask"
    out.backward(go)
    assert ext._context(q.device).last_kernel == "bwd_tc_bf16_d128_tma_mask"
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ref = ref_sdpa(torch, qr, kr, vr, mask=mask)
    ref.backward(go.float())
    assert relmax(out, ref) < 1e-2
    for a, b in ((q.grad, qr.grad), (k.grad, kr.grad), (v.grad, vr.grad)):
        assert relmax(a, b) < 2e-2


@pytest.mark.parametrize("prec,mode,min_cos", [(3, 0, 0.99), (3, 2, 0.99), (4, 2, 0.93)])
def test_quantised_autograd(torch_mod, adapter, prec, mode, min_cos):
    """Bounds: cosine >= 0.99 (int8) / 0.93 (int4 at this small size) on O vs fp32 math, and the reference's own gate
    (rel-L2 < 0.25 forward and backward, QuantizedAttentionTest.swift:519-520,651-652) for int8."""
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(9)
    q, k, v = (torch.randn(1, 2, 256, 128, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True) for _ in range(3))
    go = torch.randn(1, 2, 256, 128, device="cuda", generator=g).to(torch.bfloat16)
    out = ext.metal_quantized_flash_attention_autograd(q, k, v, False, 0.0, prec, mode)
    out.backward(go)
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ref = ref_sdpa(torch, qr, kr, vr)
    ref.backward(go.float())
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
    assert cos >= min_cos
    if prec == 3:
        rel_l2 = lambda a, b: float((a.float() - b).norm() / b.norm())
        assert rel_l2(out, ref.detach()) < 0.25
        for a, b in ((q.grad, qr.grad), (k.grad, kr.grad), (v.grad, vr.grad)):
            assert rel_l2(a, b) < 0.25


def test_quantisation_mode_routes_sdpa(torch_mod, adapter):
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(2)
    q, k, v = (torch.randn(1, 2, 256, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    ext.set_quantization_mode(ext.QUANT_INT8, ext.QUANT_BLOCK_WISE)
    ext.reset_dispatch_stats()
    try:
        with p.use_metal_sdpa():
            out = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    finally:
        ext.clear_quantization_mode()
    assert ext.get_dispatch_stats()["quantized_autograd"] == 1
    assert ext._context(q.device).last_kernel.startswith("fwd_tcq_int8_")
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref_sdpa(torch, q, k, v).flatten(), dim=0).item()
    assert cos >= 0.99


def test_rope_sdpa_and_hadamard(torch_mod, adapter):
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(4)
    B, H, S, D = 1, 2, 128, 64
    q, k, v = (torch.randn(B, H, S, D, device="cuda", generator=g) for _ in range(3))
    ang = torch.rand(S, D // 2, device="cuda", generator=g) * 6.28
    cos, sin = ang.cos().repeat_interleave(2, -1), ang.sin().repeat_interleave(2, -1)      # pair-duplicated [S, D]

    def rot(x):
        x0, x1 = x[..., 0::2], x[..., 1::2]
        c, s = ang.cos(), ang.sin()
        return torch.stack((x0 * c - x1 * s, x0 * s + x1 * c), -1).flatten(-2)
    out = ext.rope_scaled_dot_product_attention(q, k, v, cos, sin)
    assert relmax(out, ref_sdpa(torch, rot(q), rot(k), v)) < 1e-5
    x = torch.randn(8, 256, device="cuda", generator=g)
    y = ext.hadamard_rotate(x.clone(), 64)
    y2 = ext.hadamard_rotate(y.clone(), 64)             # H is an involution with the 1/sqrt(n) scaling
    assert relmax(y2, x) < 1e-5 and abs(float(y.norm() / x.norm()) - 1.0) < 1e-5


def test_rope_sdpa_gradients(torch_mod, adapter):
    """rope_scaled_dot_product_attention under autograd: dQ / dK come back through the inverse rotation (negate_sin), dV
    straight from the attention backward -- against torch autograd of the same formula, and dQ against the CPU oracle."""
    from oracle import oracle as O
    torch = torch_mod
    p, ext = adapter
    g = torch.Generator(device="cuda").manual_seed(9)
    B, H, S, D = 2, 2, 96, 64
    q, k, v = (torch.randn(B, H, S, D, device="cuda", generator=g).requires_grad_(True) for _ in range(3))
    d_o = torch.randn(B, H, S, D, device="cuda", generator=g)
    ang = torch.rand(B, S, D // 2, device="cuda", generator=g) * 6.28
    cos, sin = ang.cos().repeat_interleave(2, -1), ang.sin().repeat_interleave(2, -1)      # per-batch pair-duplicated [B, S, D]

    def rot(x):
        x0, x1 = x[..., 0::2], x[..., 1::2]
        c, s = ang.cos()[:, None], ang.sin()[:, None]
        return torch.stack((x0 * c - x1 * s, x0 * s + x1 * c), -1).flatten(-2)
    out = ext.rope_scaled_dot_product_attention(q, k, v, cos, sin, is_causal=True)
    out.backward(d_o)
    got = [t.grad.clone() for t in (q, k, v)]
    for t in (q, k, v):
        t.grad = None
    ref = ref_sdpa(torch, rot(q), rot(k), v, causal=True)
    ref.backward(d_o)
    assert relmax(out, ref) < 1e-5
    for a, t in zip(got, (q, k, v)):
        assert relmax(a, t.grad) < 1e-4
    # the oracle: rotate on the CPU, attention backward, inverse rotation of dQ
    qn, kn, vn, dn = (t.detach().cpu().numpy() for t in (q, k, v, d_o))
    cn, sn = cos.cpu().numpy(), sin.cpu().numpy()
    rq, rk, _, _ = O.attention_backward(O.rope_rotate(qn, cn, sn), O.rope_rotate(kn, cn, sn), vn, dn, causal=True)
    dq_ref = O.rope_rotate(rq, cn, sn, negate_sin=True)
    assert float(np.abs(got[0].cpu().numpy() - dq_ref).max() / np.abs(dq_ref).max()) < 1e-4


def test_unsupported_inputs_fall_back_or_raise(torch_mod, adapter):
    torch = torch_mod
    p, ext = adapter
    q = torch.randn(1, 1, 16, 16, device="cuda", dtype=torch.float64)
    ext.reset_dispatch_stats()
    with p.use_metal_sdpa():
        out = torch.nn.functional.scaled_dot_product_attention(q, q, q)          # fp64: PyTorch's own path
    assert out.dtype == torch.float64 and ext.get_dispatch_stats()["fallback_native"] == 1
    with pytest.raises(RuntimeError):
        ext.metal_scaled_dot_product_attention(q, q, q)
    with pytest.raises(RuntimeError):
        ext.metal_scaled_dot_product_attention(q.float(), q.float(), q.float(), dropout_p=0.1)


def test_matches_oracle_causal(torch_mod, adapter):
    from oracle import oracle as O
    torch = torch_mod
    p, ext = adapter
    rng = np.random.default_rng(1)
    q, k, v = (rng.standard_normal((1, 2, 192, 64)).astype(np.float32) for _ in range(3))
    ref, _ = O.attention_forward(q, k, v, causal=True)
    out = ext.metal_scaled_dot_product_attention(*(torch.from_numpy(x).cuda() for x in (q, k, v)), is_causal=True)
    assert float(np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()) < 1e-5
