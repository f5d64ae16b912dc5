"""CPU tests: the oracle against the reference's known-answer tests and the committed golden vectors."""
import numpy as np
import pytest

from oracle import oracle as O

LN2 = np.log(2.0)


# ---- quantiser KATs (metal-flash-attention/Tests/FlashAttentionTests/QuantizedAttentionTest.swift:30-59)
def test_quant_params_kat():
    x = np.array([[-10.0, -5.0, 0.0, 5.0, 10.0]], np.float32)
    _, s8 = O.quantize(x, bits=8)
    _, s4 = O.quantize(x, bits=4)
    assert abs(s8[0] - 10.0 / 127.0) < 1e-6
    assert abs(s4[0] - 10.0 / 7.0) < 1e-6
    assert s8[0] == np.float32(10.0) / np.float32(127.0)
    assert s4[0] == np.float32(10.0) / np.float32(7.0)


# ---- round trip error < 2*scale (QuantizedAttentionTest.swift:61-161)
@pytest.mark.parametrize("bits", [8, 4])
def test_quant_roundtrip_kat(bits):
    x = np.arange(-10.0, 10.0, 0.5, dtype=np.float32)[None, :]
    codes, sc = O.quantize(x, bits=bits)
    y = O.dequantize(codes, sc, 1, x.shape[1], bits=bits)
    assert np.abs(y - x).max() < 2 * sc[0]
    assert codes.nbytes == (x.size if bits == 8 else (x.size + 1) // 2)   # QuantizationTests.swift:33-45


# ---- Tests/QuantizationTests/QuantizationTests.swift:7-67
def test_quant_rmse_kats():
    x = np.arange(-5.0, 5.0, 0.1, dtype=np.float32)[None, :]
    c, s = O.quantize(x, bits=8)
    assert np.sqrt(np.mean((O.dequantize(c, s, 1, x.size, bits=8) - x) ** 2)) < 0.1
    x = np.arange(-1.0, 1.0, 0.01, dtype=np.float32)[None, :]
    c, s = O.quantize(x, bits=4)
    assert np.sqrt(np.mean((O.dequantize(c, s, 1, x.size, bits=4) - x) ** 2)) < 0.2
    z = np.zeros((1, 64), np.float32)
    c, s = O.quantize(z, bits=8)
    assert (O.dequantize(c, s, 1, 64, bits=8) == 0).all()
    k = np.full((1, 64), 5.0, np.float32)
    c, s = O.quantize(k, bits=8)
    assert np.sqrt(np.mean((O.dequantize(c, s, 1, 64, bits=8) - k) ** 2)) < 0.1


def test_int4_packing_low_nibble_first():
    # QuantizationTests.swift:95-102: byte = (hi << 4) | lo, element 2i in the low nibble, stored +8
    x = np.array([[7.0, -7.0, 3.5, 0.0, -1.0]], np.float32)
    codes, sc = O.quantize(x, bits=4)
    assert sc[0] == np.float32(1.0)
    q = [7, -7, 4, 0, -1]          # round half away from zero: 3.5 -> 4
    exp = [(q[0] + 8) | ((q[1] + 8) << 4), (q[2] + 8) | ((q[3] + 8) << 4), (q[4] + 8) | (8 << 4)]
    assert codes.tolist() == exp


def test_round_half_away_and_clamp():
    x = np.array([[0.5, -0.5, 1.5, -1.5, 2.5, 127.0, -127.0]], np.float32)
    codes, sc = O.quantize(x, bits=8)
    assert sc[0] == np.float32(1.0)
    assert codes.tolist() == [1, -1, 2, -2, 3, 127, -127]


def test_blockwise_layout_row_major_blocks():
    # GEMMQuantization.swift:567-584: blockIndex(r,c) = (r/bs)*ceil(cols/bs) + c/bs
    rng = np.random.default_rng(0)
    x = rng.standard_normal((10, 12)).astype(np.float32)
    codes, sc = O.quantize(x, bits=8, block_rows=4, block_cols=8)
    assert sc.shape == (3 * 2,)
    for br in range(3):
        for bc in range(2):
            blk = x[br * 4:(br + 1) * 4, bc * 8:(bc + 1) * 8]
            s = np.float32(np.abs(blk).max()) / np.float32(127.0)
            assert sc[br * 2 + bc] == s
            q = np.clip(np.sign(blk) * np.floor(np.abs(blk / s) + np.float32(0.5)), -128, 127)
            assert (codes.reshape(10, 12)[br * 4:(br + 1) * 4, bc * 8:(bc + 1) * 8] == q).all()


def test_gpu_variant_scale_floor():
    z = np.zeros((2, 8), np.float32)
    _, s = O.quantize(z, bits=8, clamp_scale_min=1e-8)
    assert s[0] == np.float32(1e-8)          # GEMMRuntimeQuantization.swift:89


# ---- attention: double-accumulated oracle vs the fp32 restatement of the Swift oracle (tol 2e-5,
# SquareAttentionTest.swift:557-571), on the reference's LCG inputs and ragged shapes (:6-25)
@pytest.mark.parametrize("N,D", [(10, 3), (10, 80), (8, 2), (9, 2), (23, 2), (24, 2), (25, 2), (192, 77), (192, 80),
                                 (93, 32), (99, 35), (64, 32), (32, 64), (4, 1), (4, 2), (384, 95), (777, 199)])
def test_double_oracle_matches_f32_restatement(N, D):
    q = O.lcg_ffi(42, N * D).reshape(N, D)
    k = O.lcg_ffi(43, N * D).reshape(N, D)
    v = O.lcg_ffi(44, N * D).reshape(N, D)
    o32, lse32 = O.attention_forward_f32(q, k, v)
    o64, l64 = O.attention_forward(q[None, None], k[None, None], v[None, None])
    assert np.abs(o64[0, 0] - o32).max() < 2e-5
    assert np.abs(l64[0, 0] / O.LOG2E - lse32).max() < 2e-5     # L compared after /log2e (:424-426)


GOLD_FWD = ["c1_fp32", "causal_fp32", "rect_fp32", "boolmask_fp32", "addmask_fp32", "bf16_d128", "fp16_causal_d64"]


@pytest.mark.parametrize("name", GOLD_FWD)
def test_oracle_vs_golden_forward(golden, name):
    q, k, v, ref = (golden[f"{name}.{t}"] for t in "qkvo")
    causal, scale = golden[f"{name}.meta"]
    mask = golden[f"{name}.mask"] if f"{name}.mask" in golden else None
    o, _ = O.attention_forward(q, k, v, causal=bool(causal), scale=None if scale < 0 else scale, mask=mask)
    # fp32 golden: rtol 1e-5 / atol 1e-6 (conftest.py:189-190 of the reference adapter tests)
    np.testing.assert_allclose(o, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["causal_fp32", "rect_fp32", "addmask_fp32"])
def test_oracle_vs_golden_backward(golden, name):
    q, k, v, d_o = (golden[f"{name}.{t}"] for t in ("q", "k", "v", "do"))
    causal, scale = golden[f"{name}.meta"]
    mask = golden[f"{name}.mask"] if f"{name}.mask" in golden else None
    dq, dk, dv, _ = O.attention_backward(q, k, v, d_o, causal=bool(causal), scale=None if scale < 0 else scale,
                                         mask=mask)
    for got, key in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
        np.testing.assert_allclose(got, golden[f"{name}.{key}"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", ["window_fp32", "window_bf16_d128"])
def test_oracle_sliding_window_vs_torch_golden(golden, name):
    """causal + sliding window (AttentionKernel+Softmax.swift:445,450) against torch CPU SDPA fed the equivalent banded
    bool mask (tests/golden/make_golden.py): pins the window numerics that upstream's own tests only compile."""
    q, k, v, ref = (golden[f"{name}.{t}"] for t in "qkvo")
    causal, scale, window = golden[f"{name}.meta"]
    o, _ = O.attention_forward(q, k, v, causal=bool(causal), window=int(window))
    np.testing.assert_allclose(o, ref, rtol=1e-5, atol=1e-6)
    if f"{name}.do" in golden:
        dq, dk, dv, _ = O.attention_backward(q, k, v, golden[f"{name}.do"], causal=bool(causal), window=int(window))
        for got, key in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
            np.testing.assert_allclose(got, golden[f"{name}.{key}"], rtol=1e-4, atol=1e-6)


def test_mask_rules_causal_window():
    # AttentionKernel+Softmax.swift:445,450: causal masks col > row; window masks row > col + W
    rng = np.random.default_rng(1)
    B, H, S, D, W = 1, 1, 24, 8, 5
    q, k, v = (rng.standard_normal((B, H, S, D)).astype(np.float32) for _ in range(3))
    r, c = np.arange(S)[:, None], np.arange(S)[None, :]
    keep = (c <= r) & ~(r > c + W)
    o1, l1 = O.attention_forward(q, k, v, causal=True, window=W)
    o2, l2 = O.attention_forward(q, k, v, mask=keep[None, None])
    np.testing.assert_allclose(o1, o2, rtol=0, atol=0)
    np.testing.assert_allclose(l1, l2, rtol=0, atol=0)
    assert keep.sum(axis=1).max() == W + 1


def test_side_output_conventions():
    # L = log2e * logsumexp(scale*S); D = scale * rowsum(dO*O)  
    rng = np.random.default_rng(2)
    q, k, v, d_o = (rng.standard_normal((1, 2, 16, 8)).astype(np.float32) for _ in range(4))
    scale = 0.3
    o, L = O.attention_forward(q, k, v, scale=scale)
    s = np.einsum("bhqd,bhkd->bhqk", q.astype(np.float64), k.astype(np.float64)) * scale
    lse = np.log(np.exp(s).sum(-1))
    np.testing.assert_allclose(L, lse * O.LOG2E, rtol=1e-6)
    *_, dt = O.attention_backward(q, k, v, d_o, scale=scale)
    np.testing.assert_allclose(dt, scale * (d_o * o).sum(-1), rtol=1e-5, atol=1e-6)


def test_external_mask_modes():
    rng = np.random.default_rng(3)
    q, k, v = (rng.standard_normal((1, 1, 8, 4)).astype(np.float32) for _ in range(3))
    m = rng.standard_normal((1, 1, 8, 8)).astype(np.float32)
    scale = 0.5
    o_pt, _ = O.attention_forward(q, k, v, scale=scale, mask=m, mask_mode=0)
    o_ref, _ = O.attention_forward(q, k, v, scale=scale, mask=m / scale, mask_mode=1)
    np.testing.assert_allclose(o_pt, o_ref, rtol=1e-6, atol=1e-7)   # quirk Q7: reference scales the mask too


def test_lcg_generators_known_values():
    a = O.lcg_ffi(42, 3)
    rng = 42
    exp = []
    for _ in range(3):
        rng = (rng * 1664525 + 1013904223) % (1 << 64)
        exp.append((np.float32(rng % 1000000) / np.float32(1e6) - np.float32(0.5)) * np.float32(2))
    assert a.tolist() == [float(e) for e in exp]
    b, _ = O.lcg_quantized(0x5EED5EED, 4)
    assert np.all(np.abs(b) <= 3.0) and b.std() > 0
    c = O.lcg_precision(12345, 8)
    assert c.min() >= -0.1 and c.max() <= 0.1


def test_bf16_rounding_rne():
    x = np.array([1.0, 1.00390625, 1.005859375, -2.5, 3.14159], np.float32)
    y, bits = O.round_bf16(x)
    import torch
    t = torch.from_numpy(x).to(torch.bfloat16)
    assert (t.float().numpy() == y).all()
    assert (t.view(torch.int16).numpy().view(np.uint16) == bits).all()


def test_slice_property_used_by_full_size_gpu_tests():
    """tests/test_gpu_fullsize.py checks full BASELINE sizes on slices: a row's O / L / dQ from (row, visible key span),
    a key's dK / dV from the queries that see it.  Here the oracle confirms the slices reproduce the full problem exactly."""
    rng = np.random.default_rng(0)
    B, H, N, D, W = 1, 2, 768, 32, 96
    q, k, v, g = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(4))
    of, lf = O.attention_forward(q, k, v, causal=True, window=W)
    dq, dk, dv, _ = O.attention_backward(q, k, v, g, causal=True, window=W)

    def vis(rows, keys):
        r, c = np.asarray(rows)[:, None], np.asarray(keys)[None, :]
        return (c <= r) & (r <= c + W)
    for r0 in (0, W - 20, 400, N - 48):
        rows = np.arange(r0, r0 + 48)
        keys = np.arange(max(0, r0 - W), r0 + 48)
        m = vis(rows, keys)
        o, l = O.attention_forward(q[:, :, rows], k[:, :, keys], v[:, :, keys], mask=m)
        rq, _, _, _ = O.attention_backward(q[:, :, rows], k[:, :, keys], v[:, :, keys], g[:, :, rows], mask=m)
        assert np.abs(o - of[:, :, rows]).max() < 1e-6 and np.abs(l - lf[:, :, rows]).max() < 1e-5
        assert np.abs(rq - dq[:, :, rows]).max() < 1e-6
    for j0 in (5, 300, N - 16):
        kk = np.arange(j0, j0 + 16)
        rows = np.arange(j0, min(N, j0 + 16 + W))
        keys = np.arange(max(0, rows[0] - W), rows[-1] + 1)
        _, rk, rv, _ = O.attention_backward(q[:, :, rows], k[:, :, keys], v[:, :, keys], g[:, :, rows], mask=vis(rows, keys))
        off = kk - keys[0]
        assert np.abs(rk[:, :, off] - dk[:, :, kk]).max() < 1e-6 and np.abs(rv[:, :, off] - dv[:, :, kk]).max() < 1e-6


def test_division_free_code_rule_is_bit_exact():
    """csrc/quant.cu quant_code_fast(): round(x * (1/scale)) with an exact-division fallback within 1e-4 of a half-integer
    must give the reference's codes round(x / scale) (GEMMQuantization.swift:487-521) for every input, including values that
    sit exactly on (or one ulp beside) the rounding boundaries.  Emulated here in IEEE fp32 with numpy."""
    rng = np.random.default_rng(0)
    half_away = lambda q: np.sign(q) * np.floor(np.abs(q) + np.float32(0.5))
    for trial in range(12):
        sc = np.float32(abs(rng.standard_normal()) * 10 ** rng.uniform(-4, 2) / 127)
        x = (rng.standard_normal(300_000) * sc * 60).astype(np.float32)
        edge = ((rng.integers(-130, 130, 50_000).astype(np.float32) + np.float32(0.5)) * sc).astype(np.float32)
        x = np.concatenate([x, edge, np.nextafter(edge, np.float32(0)), np.nextafter(edge, np.float32(1e9))])
        qd = (x / sc).astype(np.float32)
        exact = np.clip(half_away(qd), -128, 127)
        inv = (np.float32(1) / sc).astype(np.float32)
        t = (x * inv).astype(np.float32)
        fr = np.abs(t - np.trunc(t)).astype(np.float32)
        r = np.trunc((t + np.copysign(np.float32(0.5), t)).astype(np.float32))
        near = (np.abs(fr - np.float32(0.5)) < np.float32(1e-4)) | ~(np.abs(t) < 1024)
        r = np.clip(np.where(near, half_away(qd), r), -128, 127)
        assert np.array_equal(r, exact), trial


def test_rope_oracle_is_orthonormal_and_matches_formula():
    """oracle_rope_rotate (MFABridge.swift:269-319): formula check, shared vs per-batch tables, inverse = negate_sin."""
    rng = np.random.default_rng(0)
    B, H, S, D = 2, 3, 7, 10
    x = rng.standard_normal((B, H, S, D)).astype(np.float32)
    ang = rng.uniform(0, 6.28, (B, S, D // 2)).astype(np.float32)
    cos, sin = np.repeat(np.cos(ang), 2, -1), np.repeat(np.sin(ang), 2, -1)
    y = O.rope_rotate(x, cos, sin)
    x0, x1 = x[..., 0::2], x[..., 1::2]
    c, s = np.cos(ang)[:, None], np.sin(ang)[:, None]
    ref = np.stack((x0 * c - x1 * s, x0 * s + x1 * c), -1).reshape(x.shape)
    np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(O.rope_rotate(y, cos, sin, negate_sin=True), x, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(y, axis=-1), np.linalg.norm(x, axis=-1), rtol=1e-5)
    y1 = O.rope_rotate(x, cos[0], sin[0])                  # one [S, D] table shared by the batch
    np.testing.assert_allclose(y1[1], O.rope_rotate(x[1:], cos[0], sin[0])[0], rtol=0, atol=0)
