"""GPU parity tests for the tcgen05 backward kernels (csrc/attn_bwd_tc.cu) through the C ABI, against the CPU oracle
(oracle/attention_oracle.c: dV = P^T dO, dS = P (dP - rowsum(dO O)) scale, dQ = dS K, dK = dS^T Q).
Tolerance: 2e-2 relative to max|ref| for bf16/fp16 operands (BASELINE.json north_star).  Every case asserts the
tensor-core kernels -- not the SIMT path -- served the call."""
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
    # relative to max|ref|, floored at 1e-3 so exactly-zero gradients (single key: dS = 0) compare absolutely
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-3))


def to_dtype(x, dtype):
    if dtype == "fp16":
        h = np.asarray(x, np.float32).astype(np.float16)
        return h, h.astype(np.float32)
    vals, bits = O.round_bf16(x)
    return bits, vals


def run_case(ctx, B, H, Sq, Skv, D, dtype, causal=False, window=None, seed=0):
    import umfa
    rng = np.random.default_rng(seed)
    q, k, v, g = (rng.standard_normal(s).astype(np.float32) for s in
                  ((B, H, Sq, D), (B, H, Skv, D), (B, H, Skv, D), (B, H, Sq, D)))
    (qa, qf), (ka, kf), (va, vf), (ga, gf) = (to_dtype(x, dtype) for x in (q, k, v, g))
    w = -1 if window is None else window
    o_ref, l_ref = O.attention_forward(qf, kf, vf, causal=causal, window=w)
    dq, dk, dv, dt = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, input_precision=dtype,
                                                   causal=causal, window_size=window)
    assert ctx.last_kernel.startswith("bwd_tc_"), ctx.last_kernel
    rq, rk, rv, rt = O.attention_backward(qf, kf, vf, gf, causal=causal, window=w)
    errs = {}
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        assert np.isfinite(got).all(), name
        errs[name] = rel_max(got, ref)
    assert max(errs.values()) < 2e-2, errs
    assert rel_max(dt, rt) < 1e-3
    return errs


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("D", [128, 64])
def test_bwd_tc_square(ctx, dtype, D):
    run_case(ctx, 1, 1, 128, 128, D, dtype)
    run_case(ctx, 1, 2, 256, 256, D, dtype, seed=1)


@pytest.mark.parametrize("shape", [(1, 1, 384, 512), (2, 3, 300, 777), (1, 2, 1, 1), (1, 1, 129, 1), (1, 2, 1000, 130),
                                   (1, 1, 77, 515)])
@pytest.mark.parametrize("D", [128, 64])
def test_bwd_tc_ragged(ctx, shape, D):
    B, H, Sq, Skv = shape
    run_case(ctx, B, H, Sq, Skv, D, "bf16", seed=Sq)


@pytest.mark.parametrize("causal,window", [(True, None), (True, 100), (True, 0), (False, 64), (True, 300)])
@pytest.mark.parametrize("N", [512, 700])
def test_bwd_tc_causal_window(ctx, causal, window, N):
    run_case(ctx, 1, 2, N, N, 128, "bf16", causal=causal, window=window, seed=N)


def test_bwd_tc_causal_rect(ctx):
    # Sq != Skv with the reference's top-left aligned causal rule; rows past Skv see every key, short kv leaves
    # late key tiles without queries
    run_case(ctx, 1, 2, 640, 256, 64, "bf16", causal=True)
    run_case(ctx, 1, 2, 256, 640, 128, "fp16", causal=True)


def test_bwd_tc_long_window_property(ctx):
    """Larger shape: gradients through a sliding window equal the gradients of the equivalent dense-masked problem
    restricted to one head (oracle on one head only to keep the CPU cost bounded)."""
    errs = run_case(ctx, 1, 1, 2048, 2048, 128, "bf16", causal=True, window=512, seed=5)
    assert max(errs.values()) < 2e-2


# ---- external masks in the tensor-core backward -----------------------------------------------------------------------
def _mask_for(kind, rng, B, H, Sq, Skv):
    kw = {}
    if kind == "bool_bcast_heads":
        m = rng.random((B, 1, Sq, Skv)) > 0.4
        m[..., 0] = True
        return m, m, kw
    if kind == "bool_key_padding":
        m = np.ones((B, 1, 1, Skv), dtype=bool)
        for b in range(B):
            m[b, 0, 0, Skv - 5 - 13 * (b + 1):] = False
        return m, m, kw
    if kind == "bool_rows_empty":
        m = rng.random((B, H, Sq, Skv)) > 0.5
        m[0, H - 1, 3, :] = False
        return m, m, kw
    base = (1.5 * rng.standard_normal((B, H, Sq, Skv))).astype(np.float32)
    if kind == "add_neg_inf":
        base[rng.random(base.shape) > 0.7] = -np.inf
        base[..., 0] = 0.25
        return base, base, kw
    if kind == "add_fp16":
        m = base.astype(np.float16)
        return m, m.astype(np.float32), kw
    if kind == "add_bf16":
        om, bits = O.round_bf16(base)
        kw["mask_precision"] = "bf16"
        return bits, om, kw
    return base, base, kw


@pytest.mark.parametrize("kind", ["bool_bcast_heads", "bool_key_padding", "bool_rows_empty", "add_fp32", "add_fp16", "add_bf16", "add_neg_inf"])
@pytest.mark.parametrize("shape", [(2, 2, 256, 384, 128, "bf16"), (1, 3, 300, 333, 64, "fp16")])
def test_bwd_tc_external_masks(ctx, kind, shape):
    import umfa
    B, H, Sq, Skv, D, dtype = shape
    rng = np.random.default_rng(31)
    q, k, v, g = (rng.standard_normal(s).astype(np.float32) for s in
                  ((B, H, Sq, D), (B, H, Skv, D), (B, H, Skv, D), (B, H, Sq, D)))
    (qa, qf), (ka, kf), (va, vf), (ga, gf) = (to_dtype(x, dtype) for x in (q, k, v, g))
    m, om, kw = _mask_for(kind, rng, B, H, Sq, Skv)
    o_ref, l_ref = O.attention_forward(qf, kf, vf, mask=om)
    dq, dk, dv, dt = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, input_precision=dtype, attn_mask=m, **kw)
    assert ctx.last_kernel.startswith("bwd_tc_") and ctx.last_kernel.endswith("_mask"), ctx.last_kernel
    rq, rk, rv, rt = O.attention_backward(qf, kf, vf, gf, mask=om)
    errs = {}
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        assert np.isfinite(got).all(), name
        errs[name] = rel_max(got, ref)
    assert max(errs.values()) < 2e-2, errs


@pytest.mark.parametrize("kind", ["bool_bcast_heads", "bool_rows_empty", "add_fp16", "add_bf16"])
@pytest.mark.parametrize("shape", [(2, 2, 300, 400, 128, "bf16"), (1, 3, 520, 656, 64, "fp16"), (1, 1, 900, 1152, 128, "fp16")])
def test_bwd_tc_external_masks_staged(ctx, kind, shape):
    """dense 1- / 2-byte masks with 16-byte-multiple rows: both backward kernels read TMA-staged 128 x 128 mask tiles from shared
    memory (dQ: row owner; dK / dV: column owner); ragged query / key counts, several steps per CTA"""
    import umfa
    B, H, Sq, Skv, D, dtype = shape
    rng = np.random.default_rng(41)
    q, k, v, g = (rng.standard_normal(s).astype(np.float32) for s in
                  ((B, H, Sq, D), (B, H, Skv, D), (B, H, Skv, D), (B, H, Sq, D)))
    (qa, qf), (ka, kf), (va, vf), (ga, gf) = (to_dtype(x, dtype) for x in (q, k, v, g))
    m, om, kw = _mask_for(kind, rng, B, H, Sq, Skv)
    o_ref, l_ref = O.attention_forward(qf, kf, vf, mask=om)
    dq, dk, dv, dt = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, input_precision=dtype, attn_mask=m, **kw)
    assert ctx.last_kernel.startswith("bwd_tc_") and ctx.last_kernel.endswith("_tma_mask"), ctx.last_kernel
    rq, rk, rv, rt = O.attention_backward(qf, kf, vf, gf, mask=om)
    errs = {}
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        assert np.isfinite(got).all(), name
        errs[name] = rel_max(got, ref)
    assert max(errs.values()) < 2e-2, errs


def test_bwd_tc_staged_mask_matches_in_place_reads(ctx, monkeypatch):
    """staged tiles and in-place mask reads feed the same arithmetic: bit-identical gradients (also with causal)"""
    import umfa
    B, H, Sq, Skv, D = 1, 3, 640, 896, 128
    rng = np.random.default_rng(42)
    q, g = (rng.standard_normal((B, H, Sq, D)).astype(np.float32) for _ in range(2))
    k, v = (rng.standard_normal((B, H, Skv, D)).astype(np.float32) for _ in range(2))
    (qa, qf), (ka, kf), (va, vf), (ga, gf) = (to_dtype(x, "bf16") for x in (q, k, v, g))
    add = (2.0 * rng.standard_normal((1, H, Sq, Skv))).astype(np.float16)
    keep = rng.random((Sq, Skv)) > 0.5
    keep[:, 0] = True
    for m, causal in ((add, False), (keep, False), (keep, True)):
        o_ref, l_ref = O.attention_forward(qf, kf, vf, mask=m.astype(np.float32) if m.dtype == np.float16 else m, causal=causal)
        kw = dict(input_precision="bf16", attn_mask=m, causal=causal)
        a = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, **kw)
        assert ctx.last_kernel.endswith("_tma_mask"), ctx.last_kernel
        monkeypatch.setenv("MFA_DISABLE_MASK_TMA", "1")
        b = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, **kw)
        assert ctx.last_kernel.endswith("_mask") and not ctx.last_kernel.endswith("_tma_mask"), ctx.last_kernel
        monkeypatch.delenv("MFA_DISABLE_MASK_TMA")
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)


def test_bwd_tc_mask_with_causal_gqa_free(ctx):
    """mask + causal together, gradients against the oracle"""
    import umfa
    B, H, S, D = 1, 2, 512, 128
    rng = np.random.default_rng(32)
    q, k, v, g = (rng.standard_normal((B, H, S, D)).astype(np.float32) for _ in range(4))
    (qa, qf), (ka, kf), (va, vf), (ga, gf) = (to_dtype(x, "bf16") for x in (q, k, v, g))
    m = rng.random((S, S)) > 0.3
    m[:, 0] = True
    o_ref, l_ref = O.attention_forward(qf, kf, vf, mask=m, causal=True)
    dq, dk, dv, _ = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, input_precision="bf16", attn_mask=m, causal=True)
    assert ctx.last_kernel.endswith("_mask")
    rq, rk, rv, _ = O.attention_backward(qf, kf, vf, gf, mask=m, causal=True)
    assert max(rel_max(dq, rq), rel_max(dk, rk), rel_max(dv, rv)) < 2e-2


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("kind", ["bool", "add_fp32"])
def test_bwd_tc_mask_tile_skipping(ctx, kind, causal, monkeypatch):
    """visible-tile lists in the backward (dQ: KV tiles per query block; dK/dV: query tiles per KV tile): a packing mask with
    segments that do not line up with the tiles, some fully hidden rows; gradients against the oracle and against the
    walk-every-tile path"""
    import umfa
    B, H, S, D = 2, 2, 900, 128
    rng = np.random.default_rng(41)
    q, k, v, g = (rng.standard_normal((B, H, S, D)).astype(np.float32) for _ in range(4))
    (qa, qf), (ka, kf), (va, vf), (ga, gf) = (to_dtype(x, "bf16") for x in (q, k, v, g))
    seg = np.repeat(np.arange(6), S // 6)
    vis = np.broadcast_to((seg[:, None] == seg[None, :])[None, None], (B, 1, S, S)).copy()
    vis[1, 0, 700:, :] = False
    if causal:
        vis[1, 0, 700:, 0] = True          # keep every row alive under the causal rule for a well-defined comparison
    m = vis if kind == "bool" else np.where(vis, rng.standard_normal(vis.shape), -np.inf).astype(np.float32)
    o_ref, l_ref = O.attention_forward(qf, kf, vf, mask=m, causal=causal)
    rq, rk, rv, _ = O.attention_backward(qf, kf, vf, gf, mask=m, causal=causal)
    dq, dk, dv, _ = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, input_precision="bf16", attn_mask=m, causal=causal)
    assert ctx.last_kernel.endswith("_mask")
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        assert np.isfinite(got).all(), name
        assert rel_max(got, ref) < 2e-2, name
    monkeypatch.setenv("MFA_DISABLE_MASK_SKIP", "1")
    dq2, dk2, dv2, _ = umfa.flash_attention_backward(ctx, ga, qa, ka, va, o_ref, l_ref, input_precision="bf16", attn_mask=m, causal=causal)
    assert rel_max(dq, dq2) < 1e-5 and rel_max(dk, dk2) < 1e-5 and rel_max(dv, dv2) < 1e-5


@pytest.mark.parametrize("D", [40, 80, 96, 112])
def test_bwd_tc_padded_head_dims(ctx, D):
    """multiples of 8 other than 64 / 128 run on the next kernel width (TMA zero-fill in, clipped gradient stores out)"""
    run_case(ctx, 1, 2, 384, 512, D, "bf16", causal=True, seed=D)
    assert ctx.last_kernel.startswith("bwd_tc_"), ctx.last_kernel
