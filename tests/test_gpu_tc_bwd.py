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
