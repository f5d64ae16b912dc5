"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libMFAFFI.so), against the CPU oracle and the
committed golden vectors.  Tolerances: fp32 1e-5 relative (BASELINE.json north_star), bf16/fp16 2e-2; the six
quantities the reference checks (O, L, D, dV, dK, dQ -- SquareAttentionTest.swift:215-572)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import umfa
    c = umfa.MFAContext()
    yield c
    c.close()


def rel_max(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / (np.abs(b).max() + 1e-30))


def to_dtype(x, dtype):
    """Round fp32 data to the dtype under test; returns (array to hand to the library, fp32 values for the oracle)."""
    if dtype == "fp32":
        return np.ascontiguousarray(x, np.float32), np.asarray(x, np.float32)
    if dtype == "fp16":
        h = np.asarray(x, np.float32).astype(np.float16)
        return h, h.astype(np.float32)
    vals, bits = O.round_bf16(x)
    return bits, vals


TOL = {"fp32": 1e-5, "fp16": 2e-2, "bf16": 2e-2}


def lcg_qkv(B, H, Sq, Skv, D, seeds=(42, 43, 44)):
    q = O.lcg_ffi(seeds[0], B * H * Sq * D).reshape(B, H, Sq, D)
    k = O.lcg_ffi(seeds[1], B * H * Skv * D).reshape(B, H, Skv, D)
    v = O.lcg_ffi(seeds[2], B * H * Skv * D).reshape(B, H, Skv, D)
    return q, k, v


# ---- config 1 of BASELINE.json: fp32 non-causal B=1 H=1 N=512 D=64 (golden vector + oracle)
def test_config1_fp32_golden(ctx, golden):
    import umfa
    q, k, v, ref = (golden[f"c1_fp32.{t}"] for t in "qkvo")
    out, lse = umfa.flash_attention_forward(ctx, q, k, v, input_precision="fp32", layout="bhsd", return_lse=True)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)      # reference adapter tolerance (conftest.py:189-190)
    o_ref, l_ref = O.attention_forward(q, k, v)
    assert rel_max(out, o_ref) < 1e-5
    assert np.abs(lse - l_ref).max() < 2e-5 * O.LOG2E


GOLD = ["causal_fp32", "rect_fp32", "boolmask_fp32", "addmask_fp32", "bf16_d128", "fp16_causal_d64"]


@pytest.mark.parametrize("name", GOLD)
def test_golden_forward(ctx, golden, name):
    import umfa
    q, k, v, ref = (golden[f"{name}.{t}"] for t in "qkvo")
    causal, scale = golden[f"{name}.meta"]
    mask = golden[f"{name}.mask"] if f"{name}.mask" in golden else None
    dtype = "bf16" if name.startswith("bf16") else "fp16" if name.startswith("fp16") else "fp32"
    (qa, _), (ka, _), (va, _) = (to_dtype(x, dtype) for x in (q, k, v))
    out = umfa.flash_attention_forward(ctx, qa, ka, va, input_precision=dtype, output_precision="fp32", layout="bhsd",
                                       causal=bool(causal), softmax_scale=None if scale < 0 else float(scale),
                                       attn_mask=mask)
    assert out.dtype == np.float32
    if dtype == "fp32":
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)
    else:
        assert rel_max(out, ref) < TOL[dtype]


@pytest.mark.parametrize("name", ["window_fp32", "window_bf16_d128"])
def test_golden_sliding_window(ctx, golden, name):
    """causal + sliding window through mfa_attention_forward_ex / _backward_ex against torch CPU SDPA's banded-mask result"""
    import umfa
    q, k, v, ref = (golden[f"{name}.{t}"] for t in "qkvo")
    _, scale, window = golden[f"{name}.meta"]
    dtype = "bf16" if "bf16" in name else "fp32"
    (qa, _), (ka, _), (va, _) = (to_dtype(x, dtype) for x in (q, k, v))
    out, lse = umfa.flash_attention_forward(ctx, qa, ka, va, input_precision=dtype, output_precision="fp32", layout="bhsd",
                                            causal=True, window_size=int(window), return_lse=True)
    if dtype == "fp32":
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)
        d_o = golden[f"{name}.do"]
        dq, dk, dv, _ = umfa.flash_attention_backward(ctx, d_o, q, k, v, out, lse, input_precision="fp32", causal=True,
                                                      window_size=int(window))
        for got, key in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
            np.testing.assert_allclose(got, golden[f"{name}.{key}"], rtol=1e-4, atol=1e-6)
    else:
        assert ctx.last_kernel.startswith("fwd_tc_"), ctx.last_kernel
        assert rel_max(out, ref) < TOL[dtype]


# ---- the reference's 20 ragged (N, D) shapes (SquareAttentionTest.swift:6-25), all six quantities, fp32
SHAPES = [(10, 3), (10, 80), (8, 2), (9, 2), (23, 2), (24, 2), (25, 2), (192, 77), (192, 80), (93, 32), (99, 35),
          (64, 32), (32, 64), (4, 1), (4, 2), (384, 95), (777, 199), (256, 128), (512, 256), (1, 1)]


@pytest.mark.parametrize("N,D", SHAPES)
def test_ragged_shapes_six_quantities_fp32(ctx, N, D):
    import umfa
    q, k, v = lcg_qkv(1, 1, N, N, D)
    d_o = O.lcg_ffi(45, N * D).reshape(1, 1, N, D)
    out, lse = umfa.flash_attention_forward(ctx, q, k, v, input_precision="fp32", layout="bhsd", return_lse=True)
    o_ref, l_ref = O.attention_forward(q, k, v)
    tol = 2e-5                                              # SquareAttentionTest.swift:565-570
    assert np.abs(out - o_ref).max() < tol
    assert np.abs(lse - l_ref).max() / O.LOG2E < tol        # L compared after /log2e (:424-426)
    dq, dk, dv, dt = umfa.flash_attention_backward(ctx, d_o, q, k, v, out, lse, input_precision="fp32")
    rq, rk, rv, rt = O.attention_backward(q, k, v, d_o)
    scale = 1 / np.sqrt(D)
    assert np.abs(dt - rt).max() / scale < tol * max(1.0, np.abs(rt).max() / scale)   # D compared after /scale (:427-429)
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        assert np.abs(got - ref).max() < tol * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dtype", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("B,H,Sq,Skv,D,causal,window", [
    (2, 3, 65, 65, 64, False, None), (1, 2, 200, 200, 128, True, None), (1, 2, 96, 160, 32, False, None),
    (2, 2, 150, 150, 64, True, 17), (1, 1, 130, 130, 128, False, 40), (1, 4, 33, 257, 80, False, None),
])
def test_forward_backward_multihead(ctx, dtype, B, H, Sq, Skv, D, causal, window):
    import umfa
    rng = np.random.default_rng(B * 1000 + Sq + D)
    q, k, v, d_o = (rng.standard_normal(s).astype(np.float32) for s in
                    ((B, H, Sq, D), (B, H, Skv, D), (B, H, Skv, D), (B, H, Sq, D)))
    (qa, qf), (ka, kf), (va, vf), (ga, gf) = (to_dtype(x, dtype) for x in (q, k, v, d_o))
    w = -1 if window is None else window
    out, lse = umfa.flash_attention_forward(ctx, qa, ka, va, input_precision=dtype, output_precision="fp32",
                                            layout="bhsd", causal=causal, window_size=window, return_lse=True)
    o_ref, l_ref = O.attention_forward(qf, kf, vf, causal=causal, window=w)
    assert rel_max(out, o_ref) < TOL[dtype]
    assert np.abs(lse - l_ref).max() < (1e-4 if dtype == "fp32" else 5e-2)
    dq, dk, dv, dt = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, input_precision=dtype,
                                                   causal=causal, window_size=window)
    rq, rk, rv, rt = O.attention_backward(qf, kf, vf, gf, causal=causal, window=w)
    btol = 1e-4 if dtype == "fp32" else TOL[dtype]
    for got, ref in ((dq, rq), (dk, rk), (dv, rv), (dt, rt)):
        assert rel_max(got, ref) < btol


def test_backward_golden(ctx, golden):
    import umfa
    for name in ("causal_fp32", "rect_fp32", "addmask_fp32"):
        q, k, v, d_o = (golden[f"{name}.{t}"] for t in ("q", "k", "v", "do"))
        causal, scale = golden[f"{name}.meta"]
        mask = golden[f"{name}.mask"] if f"{name}.mask" in golden else None
        sc = None if scale < 0 else float(scale)
        out, lse = umfa.flash_attention_forward(ctx, q, k, v, input_precision="fp32", layout="bhsd", causal=bool(causal),
                                                softmax_scale=sc, attn_mask=mask, return_lse=True)
        dq, dk, dv, _ = umfa.flash_attention_backward(ctx, d_o, q, k, v, out, lse, input_precision="fp32",
                                                      causal=bool(causal), softmax_scale=sc, attn_mask=mask)
        for got, key in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
            np.testing.assert_allclose(got, golden[f"{name}.{key}"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("mask_kind", ["bool2d", "bool4d_bcast_heads", "add_fp32", "add_fp16", "add_bf16", "bool_rows_empty"])
def test_external_masks(ctx, mask_kind):
    import umfa
    B, H, Sq, Skv, D = 2, 3, 70, 90, 64
    rng = np.random.default_rng(5)
    q, k, v = (rng.standard_normal(s).astype(np.float32) for s in ((B, H, Sq, D), (B, H, Skv, D), (B, H, Skv, D)))
    kw = {}
    if mask_kind == "bool2d":
        m = rng.random((Sq, Skv)) > 0.4
        m[:, 0] = True
        om = m
    elif mask_kind == "bool4d_bcast_heads":
        m = rng.random((B, 1, Sq, Skv)) > 0.5
        m[..., 3] = True
        om = m
    elif mask_kind == "bool_rows_empty":
        m = rng.random((B, H, Sq, Skv)) > 0.5
        m[0, 1, 5, :] = False            # a fully masked row: O = 0, L = -inf by contract
        om = m
    else:
        base = rng.standard_normal((B, H, Sq, Skv)).astype(np.float32)
        if mask_kind == "add_fp16":
            m = base.astype(np.float16); om = m.astype(np.float32)
        elif mask_kind == "add_bf16":
            om, bits = O.round_bf16(base); m = bits; kw["mask_precision"] = "bf16"
        else:
            m = base; om = base
    out, lse = umfa.flash_attention_forward(ctx, q, k, v, input_precision="fp32", layout="bhsd", attn_mask=m,
                                            return_lse=True, **kw)
    o_ref, l_ref = O.attention_forward(q, k, v, mask=om)
    assert rel_max(out, o_ref) < 1e-5
    fin = np.isfinite(l_ref)
    assert np.array_equal(np.isfinite(lse), fin)
    assert np.abs(lse[fin] - l_ref[fin]).max() < 1e-4
    if mask_kind == "bool_rows_empty":
        assert (out[0, 1, 5] == 0).all() and lse[0, 1, 5] == -np.inf


def test_reference_python_adapter_layout_bshd(ctx):
    """4-D arrays in the reference adapter's [batch, seq, heads, dim] layout go through BHSD strides."""
    import umfa
    rng = np.random.default_rng(9)
    B, S, H, D = 2, 50, 3, 32
    q, k, v = (rng.standard_normal((B, S, H, D)).astype(np.float32) for _ in range(3))
    out = umfa.flash_attention_forward(ctx, q, k, v, input_precision="fp32", causal=True)
    o_ref, _ = O.attention_forward(*(x.transpose(0, 2, 1, 3) for x in (q, k, v)), causal=True)
    assert out.shape == q.shape
    assert rel_max(out.transpose(0, 2, 1, 3), o_ref) < 1e-5


def test_ffi_smoke_shapes_scales_patterns(ctx):
    """Tests/MFAFFITests/MFAFFITests.swift: minimal sizes 1x1/1x4/2x2, scale sweep 0.01..100, finite outputs."""
    import umfa
    for (S, D) in ((1, 1), (1, 4), (2, 2), (128, 16), (256, 64)):
        q, k, v = (O.lcg_ffi(s, S * D).reshape(S, D) for s in (12345, 12346, 12347))
        for scale in (0.01, 0.1, 1.0, 10.0, 100.0):
            out = umfa.flash_attention_forward(ctx, q, k, v, input_precision="fp32", softmax_scale=scale)
            ref, _ = O.attention_forward(q[None, None], k[None, None], v[None, None], scale=scale)
            assert np.isfinite(out).all()
            # the upstream sweep only asserts finiteness (MFAFFITests.swift:295-313); at scale >= 10 the logits reach
            # |z| ~ 1e3 where one fp32 ulp of z already moves exp(z) by ~1e-5, so the 1e-5 gate applies to scale <= 1
            assert rel_max(out, ref[0, 0]) < (1e-5 if scale <= 1.0 else 2e-4)


def test_fp16_output_written_in_place_like_reference_adapter(ctx):
    import umfa
    rng = np.random.default_rng(11)
    q, k, v = (rng.standard_normal((64, 32)).astype(np.float16) for _ in range(3))
    out = umfa.flash_attention_forward(ctx, q, k, v)            # adapter defaults: fp16 in / fp16 out
    assert out.dtype == np.float16 and out.shape == q.shape
    ref, _ = O.attention_forward(*(x.astype(np.float32)[None, None] for x in (q, k, v)))
    assert rel_max(out.astype(np.float32), ref[0, 0]) < 2e-2


def test_zero_copy_aliasing_and_buffer_api(ctx):
    """Outputs appear in the caller's array (MultiHeadFFITests.swift:1008-1014); create_buffer contents are CPU-visible."""
    import ctypes
    import umfa
    from umfa._ffi import _lib
    S, D = 32, 16
    q, k, v = (O.lcg_ffi(s, S * D).reshape(S, D) for s in (100, 101, 102))
    out = np.full((S, D), np.nan, np.float32)
    bufs = [umfa.MFABuffer(ctx, a) for a in (q, k, v, out)]
    rc = _lib.mfa_attention_forward(ctx.handle, *[b.handle for b in bufs], 1, S, S, 1, D, 0.25, False, 2, 2, 2,
                                    False, False, False, False, None, 0, None, None, 0, 0, 0)
    assert rc == 0 and np.isfinite(out).all()
    ref, _ = O.attention_forward(q[None, None], k[None, None], v[None, None], scale=0.25)
    assert rel_max(out, ref[0, 0]) < 1e-5
    assert ctx.gpu_latency > 0
    for b in bufs:
        b.close()
    owned = umfa.MFABuffer(ctx, size=S * D * 4)
    p = owned.contents_ptr()
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), (S, D))
    assert (arr == 0).all()
    arr[:] = q
    o2 = umfa.MFABuffer(ctx, size=S * D * 4)
    rc = _lib.mfa_attention_forward(ctx.handle, owned.handle, bufs_k(ctx, k).handle, bufs_k(ctx, v).handle, o2.handle,
                                    1, S, S, 1, D, 0.25, False, 2, 2, 2, False, False, False, False,
                                    None, 0, None, None, 0, 0, 0)
    assert rc == 0
    got = np.ctypeslib.as_array(ctypes.cast(o2.contents_ptr(), ctypes.POINTER(ctypes.c_float)), (S, D))
    assert rel_max(got, ref[0, 0]) < 1e-5


_keep = []


def bufs_k(ctx, a):
    import umfa
    b = umfa.MFABuffer(ctx, np.ascontiguousarray(a))
    _keep.append(b)
    return b


def test_context_create_destroy_loops():
    import umfa
    for _ in range(5):                       # examples/python-ffi/tests/test_basic.py:229-233
        with umfa.MFAContext() as c:
            assert c
    a, b = umfa.MFAContext(), umfa.MFAContext()
    assert a.handle.value == b.handle.value  # process-wide singleton (MFABridge.swift:782-798)
    a.close(); b.close()


def test_invalid_sizes_rejected(ctx):
    import umfa
    q = np.zeros((8, 4), np.float32)
    small = np.zeros((2, 4), np.float32)
    with pytest.raises(umfa.MFAError) as e:
        from umfa._ffi import _lib
        bufs = [umfa.MFABuffer(ctx, a) for a in (q, small, q, q)]
        umfa._ffi._check_error(_lib.mfa_attention_forward(ctx.handle, *[b.handle for b in bufs], 1, 8, 8, 1, 4, 1.0,
                                                          False, 2, 2, 2, False, False, False, False, None, 0, None,
                                                          None, 0, 0, 0))
    assert e.value.code == 1


def test_empty_inputs(ctx):
    import umfa
    q = np.zeros((0, 8), np.float32)
    kv = np.ones((4, 8), np.float32)
    assert umfa.flash_attention_forward(ctx, q, kv, kv, input_precision="fp32").shape == (0, 8)
    q = np.ones((3, 8), np.float32)
    kv = np.zeros((0, 8), np.float32)
    out, lse = umfa.flash_attention_forward(ctx, q, kv, kv, input_precision="fp32", return_lse=True)
    assert (out == 0).all() and np.isneginf(lse).all()


def test_transposed_operands(ctx):
    """transpose_x = operand stored [D, S] per head (AttentionKernel.swift:299-313)."""
    import umfa
    from umfa._ffi import _lib
    S, D = 48, 24
    q, k, v = (O.lcg_ffi(s, S * D).reshape(S, D) for s in (7777, 7778, 7779))
    out = np.zeros((D, S), np.float32)
    arrs = [np.ascontiguousarray(q.T), np.ascontiguousarray(k.T), np.ascontiguousarray(v.T), out]
    bufs = [umfa.MFABuffer(ctx, a) for a in arrs]
    rc = _lib.mfa_attention_forward(ctx.handle, *[b.handle for b in bufs], 1, S, S, 1, D, 0.2, True, 2, 2, 2,
                                    True, True, True, True, None, 0, None, None, 0, 0, 0)
    assert rc == 0
    ref, _ = O.attention_forward(q[None, None], k[None, None], v[None, None], scale=0.2, causal=True)
    assert rel_max(out.T, ref[0, 0]) < 1e-5
