"""Parity at BASELINE.json's full sizes through slice properties (the oracle is O(N^2) on the CPU, so it is run on the
sub-problems that determine sampled outputs exactly):

  * a query row's output / L / dQ depend only on that row and on the keys it sees -> oracle on (rows, visible key span)
    with an explicit bool mask built from absolute indices;
  * a key's dK / dV depend only on the queries that see it -> oracle on (queries in [j, j + W], their whole key spans).

config 4: bf16 B=1 N=32768 D=128 causal + window 4096, forward + backward (two of the 32 heads: heads are independent).
config 5: bf16 N=131072 causal, ranks 0 and 7 of an 8-way ring emulated on one GPU (one head of the 32)."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    import umfa
    c = umfa.MFAContext()
    yield c
    c.close()


def _vis(rows, keys, causal, window):
    r, c = np.asarray(rows)[:, None], np.asarray(keys)[None, :]
    m = np.ones((len(rows), len(keys)), bool)
    if causal:
        m &= c <= r
    if window is not None:
        m &= r <= c + window
    return m


def test_config4_full_size_forward_backward(ctx):
    import umfa
    B, H, N, D, W = 1, 2, 32768, 128, 4096
    rng = np.random.default_rng(40)
    q, k, v, g = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(4))
    (qf, qb), (kf, kb), (vf, vb), (gf, gb) = (O.round_bf16(x) for x in (q, k, v, g))
    out, lse = umfa.flash_attention_forward(ctx, qb, kb, vb, input_precision="bf16", output_precision="fp32", layout="bhsd",
                                            causal=True, window_size=W, return_lse=True)
    assert ctx.last_kernel.startswith("fwd_tc_")
    dq, dk, dv, _ = umfa.flash_attention_backward(ctx, gb, qb, kb, vb, out, lse, input_precision="bf16", causal=True, window_size=W)
    assert ctx.last_kernel.startswith("bwd_tc_")
    assert np.isfinite(out).all() and np.isfinite(dq).all() and np.isfinite(dk).all() and np.isfinite(dv).all()
    # rows: first tile, a window edge, the middle, the last rows
    for r0 in (0, W - 20, 17000, N - 48):
        rows = np.arange(r0, r0 + 48)
        keys = np.arange(max(0, r0 - W), r0 + 48)
        m = _vis(rows, keys, True, W)
        sl = lambda x, idx: x[:, :, idx]
        o_ref, l_ref = O.attention_forward(sl(qf, rows), sl(kf, keys), sl(vf, keys), mask=m)
        assert np.abs(out[:, :, rows] - o_ref).max() / np.abs(o_ref).max() < 2e-2, r0
        assert np.abs(lse[:, :, rows] - l_ref).max() < 2e-2, r0
        rq, _, _, _ = O.attention_backward(sl(qf, rows), sl(kf, keys), sl(vf, keys), sl(gf, rows), mask=m)
        assert np.abs(dq[:, :, rows] - rq).max() / max(np.abs(rq).max(), 1e-3) < 2e-2, r0
    # keys: dK / dV of 16 keys need every query in [j0, j0 + 15 + W]
    for j0 in (5, 20000, N - 16):
        kk = np.arange(j0, j0 + 16)
        rows = np.arange(j0, min(N, j0 + 16 + W))
        keys = np.arange(max(0, rows[0] - W), rows[-1] + 1)
        m = _vis(rows, keys, True, W)
        sl = lambda x, idx: x[:, :, idx]
        _, rk, rv, _ = O.attention_backward(sl(qf, rows), sl(kf, keys), sl(vf, keys), sl(gf, rows), mask=m)
        off = kk - keys[0]
        assert np.abs(dk[:, :, kk] - rk[:, :, off]).max() / max(np.abs(rk[:, :, off]).max(), 1e-3) < 2e-2, j0
        assert np.abs(dv[:, :, kk] - rv[:, :, off]).max() / max(np.abs(rv[:, :, off]).max(), 1e-3) < 2e-2, j0


def test_config5_full_length_ring_ranks(ctx):
    """ranks 0 and 7 of the 8-way zig-zag ring over 131072 tokens (one head), all 8 steps each, sampled rows vs the oracle"""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    from umfa import ring
    world, N, H, D = 8, 131072, 1, 128
    C = N // (2 * world)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(50)
    q, k, v = (O.round_bf16(rng.standard_normal((1, H, N, D)).astype(np.float32))[0] for _ in range(3))
    tod = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(torch.bfloat16).contiguous()
    shard = lambda x, r: tuple(tod(c) for c in ring.shard_sequence(x, r, world))

    class Emulated(ring.CudaBackend):
        def __init__(self, rank):
            super().__init__(ctx, None, dev, "bf16")
            self.rank = rank

        def exchange_start(self, buf, dst, src, step):
            s = (self.rank - step - 1) % world
            return torch.stack([self.cat_seq(shard(k, s)), self.cat_seq(shard(v, s))]).contiguous()

        def exchange_finish(self, handle, step):
            return handle

    for r in (0, world - 1):
        be = Emulated(r)
        (o_lo, l_lo), (o_hi, l_hi) = ring.ring_attention_forward(be, shard(q, r), shard(k, r), shard(v, r), r, world, 1.0 / np.sqrt(D))
        torch.cuda.synchronize(dev)
        lo, hi = ring.chunk_ids(r, world)
        for (o, l), cid in (((o_lo, l_lo), lo), ((o_hi, l_hi), hi)):
            o, l = o.cpu().numpy(), l.cpu().numpy()
            assert np.isfinite(o).all()
            for loc in (0, C // 2 + 3, C - 32):
                rows = np.arange(cid * C + loc, cid * C + loc + 32)
                keys = np.arange(0, rows[-1] + 1)
                o_ref, l_ref = O.attention_forward(q[:, :, rows], k[:, :, keys], v[:, :, keys], mask=_vis(rows, keys, True, None))
                assert np.abs(o[:, :, loc:loc + 32] - o_ref).max() / np.abs(o_ref).max() < 2e-2, (r, cid, loc)
                assert np.abs(l[:, :, loc:loc + 32] - l_ref).max() < 2e-2, (r, cid, loc)
