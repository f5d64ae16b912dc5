"""ctypes front-end of the CPU oracle (oracle/attention_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of attention_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module, and only as the checker.

Also holds the deterministic input generators of the reference's own tests so that parity tests feed
the very same numbers the upstream XCTest suites do:
  lcg_ffi          Tests/MFAFFITests/MultiHeadFFITests.swift:1533-1541   rng = rng*1664525 + 1013904223 (UInt64 wrap)
  lcg_precision    Tests/MFAFFITests/SimplePrecisionTests.swift:13-28    (& 0xFFFFFFFF) variant, /5e6 - 0.1
  lcg_quantized    metal-flash-attention/Tests/FlashAttentionTests/QuantizedAttentionTest.swift:446-453
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
LOG2E = 1.4426950408889634


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "attention_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, env={**os.environ, "CC": "", "MAKEFLAGS": ""})
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        i, f, l = ctypes.c_int, ctypes.c_float, ctypes.c_long
        L.oracle_attention_forward.argtypes = [fp, fp, fp, fp, fp, fp, i, i, i, i, i, f, i, i, i]
        L.oracle_attention_forward.restype = i
        L.oracle_attention_backward.argtypes = [fp] * 9 + [i, i, i, i, i, f, i, i, i]
        L.oracle_attention_backward.restype = i
        L.oracle_attention_forward_f32.argtypes = [fp, fp, fp, fp, fp, i, i, i]
        L.oracle_attention_forward_f32.restype = i
        L.oracle_quant_num_blocks.argtypes = [l, l, l, l]
        L.oracle_quant_num_blocks.restype = l
        u8 = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_quantize.argtypes = [fp, l, l, l, l, i, f, u8, fp]
        L.oracle_quantize.restype = i
        L.oracle_dequantize.argtypes = [u8, fp, l, l, l, l, i, fp]
        L.oracle_dequantize.restype = i
        L.oracle_round_bf16.argtypes = [fp, fp, ctypes.POINTER(ctypes.c_uint16), l]
        L.oracle_round_bf16.restype = None
        L.oracle_num_threads.restype = i
        L.oracle_rope_rotate.argtypes = [fp, fp, fp, fp, l, l, l, l, l, i]
        L.oracle_rope_rotate.restype = None
        L.oracle_set_num_threads.argtypes = [i]
        L.oracle_set_num_threads.restype = None
        _lib = L
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f32c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP thread count of the oracle (bench.py's reference arm: all host cores, whatever OMP_NUM_THREADS says)."""
    lib().oracle_set_num_threads(int(n))
    return num_threads()


def attention_forward(q, k, v, *, scale=None, causal=False, window=-1, mask=None, mask_mode=0):
    """q [B,H,Sq,D], k/v [B,H,Skv,D] float32 (already rounded to the dtype under test).
    mask: None or array broadcastable to [B,H,Sq,Skv]; bool (True = attend) or additive float.
    Returns (O float32 [B,H,Sq,D], L float32 [B,H,Sq] in log2 units: L = log2e * logsumexp)."""
    q, k, v = _f32c(q), _f32c(k), _f32c(v)
    B, H, Sq, D = q.shape
    Skv = k.shape[2]
    if scale is None:
        scale = 1.0 / np.sqrt(D)
    m = dense_mask(mask, B, H, Sq, Skv)
    o = np.empty((B, H, Sq, D), np.float32)
    lse = np.empty((B, H, Sq), np.float32)
    rc = lib().oracle_attention_forward(_fp(q), _fp(k), _fp(v), _fp(m), _fp(o), _fp(lse), B, H, Sq, Skv, D,
                                        float(scale), int(causal), int(window), int(mask_mode))
    assert rc == 0
    return o, lse


def attention_backward(q, k, v, d_o, *, scale=None, causal=False, window=-1, mask=None, mask_mode=0):
    """Returns (dQ, dK, dV, Dterm) float32; Dterm = scale * rowsum(dO*O) (reference convention)."""
    q, k, v, d_o = _f32c(q), _f32c(k), _f32c(v), _f32c(d_o)
    B, H, Sq, D = q.shape
    Skv = k.shape[2]
    if scale is None:
        scale = 1.0 / np.sqrt(D)
    m = dense_mask(mask, B, H, Sq, Skv)
    dq = np.empty_like(q)
    dk = np.empty_like(k)
    dv = np.empty_like(v)
    dt = np.empty((B, H, Sq), np.float32)
    rc = lib().oracle_attention_backward(_fp(q), _fp(k), _fp(v), _fp(m), _fp(d_o), _fp(dq), _fp(dk), _fp(dv),
                                         _fp(dt), B, H, Sq, Skv, D, float(scale), int(causal), int(window),
                                         int(mask_mode))
    assert rc == 0
    return dq, dk, dv, dt


def rope_rotate(x, cos_t, sin_t, negate_sin=False):
    """Interleaved-pair RoPE of x [B,H,S,D] with pair-duplicated fp32 tables [S,D] or [B,S,D] (MFABridge.swift:269-319)."""
    x = _f32c(x)
    B, H, S, D = x.shape
    cos_t, sin_t = _f32c(cos_t), _f32c(sin_t)
    stride = S * D if cos_t.ndim == 3 and cos_t.shape[0] == B and B > 1 else 0
    out = np.empty_like(x)
    lib().oracle_rope_rotate(_fp(x), _fp(out), _fp(cos_t), _fp(sin_t), B, H, S, D, stride, int(negate_sin))
    return out


def attention_forward_f32(q, k, v):
    """Single-head fp32 restatement of the Swift oracle: q [Sq,D], k/v [Skv,D]; scale = 1/sqrt(D).
    Returns (O, natural-log LSE)."""
    q, k, v = _f32c(q), _f32c(k), _f32c(v)
    Sq, D = q.shape
    Skv = k.shape[0]
    o = np.empty((Sq, D), np.float32)
    lse = np.empty((Sq,), np.float32)
    lib().oracle_attention_forward_f32(_fp(q), _fp(k), _fp(v), _fp(o), _fp(lse), Sq, Skv, D)
    return o, lse


def dense_mask(mask, B, H, Sq, Skv):
    """Expand a bool (True/non-zero = attend, MFABridge.swift:201-205) or additive mask to dense fp32
    [B,H,Sq,Skv] exactly as mfa_prepare_mask does (MFABridge.swift:157-242)."""
    if mask is None:
        return None
    m = np.asarray(mask)
    if m.dtype == np.bool_ or m.dtype == np.uint8:
        m = np.where(m.astype(bool), np.float32(0), np.float32(-np.inf))
    m = np.broadcast_to(m.astype(np.float32), (B, H, Sq, Skv))
    return np.ascontiguousarray(m)


# ------------------------------------------------------------------------------------------ quantiser
def quantize(x, *, bits=8, block_rows=None, block_cols=None, clamp_scale_min=0.0):
    """x: 2-D float32 [rows, cols].  Returns (codes uint8/int8 array, scales float32 [n_blocks])."""
    x = _f32c(x)
    rows, cols = x.shape
    br = rows if block_rows is None else int(block_rows)
    bc = cols if block_cols is None else int(block_cols)
    nb = lib().oracle_quant_num_blocks(rows, cols, br, bc)
    n = rows * cols
    codes = np.zeros(n if bits == 8 else (n + 1) // 2, np.uint8)
    scales = np.zeros(nb, np.float32)
    rc = lib().oracle_quantize(_fp(x), rows, cols, br, bc, bits, float(clamp_scale_min),
                               codes.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _fp(scales))
    assert rc == 0
    return (codes.view(np.int8) if bits == 8 else codes), scales


def dequantize(codes, scales, rows, cols, *, bits=8, block_rows=None, block_cols=None):
    br = rows if block_rows is None else int(block_rows)
    bc = cols if block_cols is None else int(block_cols)
    codes = np.ascontiguousarray(codes).view(np.uint8)
    scales = _f32c(scales)
    out = np.empty((rows, cols), np.float32)
    rc = lib().oracle_dequantize(codes.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _fp(scales), rows, cols,
                                 br, bc, bits, _fp(out))
    assert rc == 0
    return out


def round_bf16(x):
    """Round-to-nearest-even to bf16; returns (float32 values, uint16 bit patterns)."""
    x = _f32c(x)
    out = np.empty_like(x)
    bits = np.empty(x.shape, np.uint16)
    lib().oracle_round_bf16(_fp(x), _fp(out), bits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), x.size)
    return out, bits


def round_fp16(x):
    h = np.asarray(x, np.float32).astype(np.float16)
    return h.astype(np.float32), h.view(np.uint16)


# ------------------------------------------------------------------------ reference input generators
_M64 = (1 << 64) - 1


def lcg_ffi(seed: int, n: int) -> np.ndarray:
    """MultiHeadFFITests.swift:1533-1541: rng = rng &* 1664525 &+ 1013904223 on UInt64;
    value = (Float(rng % 1000000) / 1000000.0 - 0.5) * 2.0."""
    out = np.empty(n, np.float32)
    rng = seed & _M64
    for i in range(n):
        rng = (rng * 1664525 + 1013904223) & _M64
        out[i] = (np.float32(rng % 1000000) / np.float32(1000000.0) - np.float32(0.5)) * np.float32(2.0)
    return out


def lcg_precision(seed: int, n: int) -> np.ndarray:
    """SimplePrecisionTests.swift:13-28: rng = (rng*1664525 + 1013904223) & 0xFFFFFFFF;
    value = Float(rng % 1000000) / 5000000.0 - 0.1."""
    out = np.empty(n, np.float32)
    rng = seed & 0xFFFFFFFF
    for i in range(n):
        rng = (rng * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = np.float32(rng % 1000000) / np.float32(5000000.0) - np.float32(0.1)
    return out


def lcg_quantized(seed: int, n: int):
    """QuantizedAttentionTest.swift:446-453: 64-bit LCG, Float(Int32(truncating: seed)) / Float(Int32.max),
    then * 2 - 1.  Returns (values, next_seed) so Q, K, V can be drawn in sequence."""
    out = np.empty(n, np.float32)
    s = seed & _M64
    for i in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) & _M64
        lo = s & 0xFFFFFFFF
        i32 = lo - (1 << 32) if lo >= (1 << 31) else lo
        out[i] = np.float32(i32) / np.float32(2147483647) * np.float32(2.0) - np.float32(1.0)
    return out, s


def rel_l2(candidate, reference) -> float:
    """QuantizedAttentionTest.swift:800-816."""
    c = np.asarray(candidate, np.float64).ravel()
    r = np.asarray(reference, np.float64).ravel()
    return float(np.sqrt(((c - r) ** 2).sum()) / (np.sqrt((r ** 2).sum()) + 1e-8))
