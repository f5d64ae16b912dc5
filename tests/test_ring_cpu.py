"""CPU tests of the multi-GPU host logic (umfa/ring.py): zig-zag partition, per-step schedule, partial merge, and the
full ring driver run as a world_size-2 (and 4) torch.distributed job over gloo with the CPU oracle standing in for
the attention kernel.  The result must equal single-device causal attention from the same oracle."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
from umfa import ring  # noqa: E402


def test_chunk_ownership_covers_sequence():
    for G in (1, 2, 4, 8):
        owned = sorted(c for r in range(G) for c in ring.chunk_ids(r, G))
        assert owned == list(range(2 * G))


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_schedule_is_exact_and_balanced(G):
    """Union over ranks and steps of the scheduled chunk pairs = the lower triangle (incl. diagonal) of the 2G x 2G
    chunk grid, each pair exactly once; every rank does the same amount of work at every step."""
    seen = {}
    for r in range(G):
        for s in range(G):
            src = (r - s) % G
            qc, kc = ring.chunk_ids(r, G), ring.chunk_ids(src, G)
            cost = 0.0
            for qi, ki, causal in ring.step_schedule(r, G, s):
                pair = (qc[qi], kc[ki])
                assert pair not in seen
                seen[pair] = causal
                cost += 0.5 if causal else 1.0
            assert cost == (1.0 if G == 1 else 2.0) or (s == 0 and cost == 2.0)
    want = {(a, b): (a == b) for a in range(2 * G) for b in range(2 * G) if b <= a}
    assert seen == want


@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
def test_fused_step_plan_covers_the_schedule(G):
    """The one-launch-per-step plan (rectangular problem over the [low | high] layout) does exactly the chunk pairs of
    step_schedule, with the diagonal pairs -- and only those -- handled by the causal rule on local indices."""
    for r in range(G):
        for s in range(G):
            (q0, qn), (k0, kn), causal = ring.step_plan(r, G, s)
            got = set()
            for qi in range(q0, q0 + qn):
                for ki in range(k0, k0 + kn):
                    if causal:
                        if ki <= qi:
                            got.add((qi, ki, qi == ki))
                    else:
                        got.add((qi, ki, False))
            assert got == set(ring.step_schedule(r, G, s)), (G, r, s)


def test_merge_matches_joint_softmax():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((1, 2, 40, 16)).astype(np.float32)
    k = rng.standard_normal((1, 2, 96, 16)).astype(np.float32)
    v = rng.standard_normal((1, 2, 96, 16)).astype(np.float32)
    o_ref, l_ref = O.attention_forward(q, k, v)
    o1, l1 = O.attention_forward(q, k[:, :, :50], v[:, :, :50])
    o2, l2 = O.attention_forward(q, k[:, :, 50:], v[:, :, 50:])
    o1, l1 = np.array(o1), np.array(l1)
    ring.merge_partials_numpy(o1, l1, np.array(o2), np.array(l2))
    np.testing.assert_allclose(o1, o_ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(l1, l_ref, rtol=2e-5, atol=2e-5)
    # a partial with no visible keys (-inf) leaves the accumulator untouched
    o3, l3 = np.zeros_like(o1), np.full_like(l1, -np.inf)
    keep_o, keep_l = o1.copy(), l1.copy()
    ring.merge_partials_numpy(o1, l1, o3, l3)
    np.testing.assert_array_equal(o1, keep_o)
    np.testing.assert_array_equal(l1, keep_l)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)                      # same full problem on every rank
        B, H, D = 1, 2, 32
        q, k, v = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(3))
        scale = 1.0 / np.sqrt(D)

        def attend(qq, kk, vv, causal, sc):
            return O.attention_forward(np.ascontiguousarray(qq), np.ascontiguousarray(kk), np.ascontiguousarray(vv),
                                       causal=causal, scale=sc)
        be = ring.HostBackend(attend, dist)
        qp, kp, vp = (ring.shard_sequence(x, rank, world) for x in (q, k, v))
        (o_lo, l_lo), (o_hi, l_hi) = ring.ring_attention_forward(be, qp, kp, vp, rank, world, scale)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), o_lo=o_lo, l_lo=l_lo, o_hi=o_hi, l_hi=l_hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 96), (4, 128)])
def test_ring_attention_gloo_matches_single_device(tmp_path, world, N):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    B, H, D = 1, 2, 32
    q, k, v = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(3))
    o_ref, l_ref = O.attention_forward(q, k, v, causal=True)
    c = N // (2 * world)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        lo, hi = ring.chunk_ids(r, world)
        np.testing.assert_allclose(got["o_lo"], o_ref[:, :, lo * c:(lo + 1) * c], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(got["o_hi"], o_ref[:, :, hi * c:(hi + 1) * c], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(got["l_lo"], l_ref[:, :, lo * c:(lo + 1) * c], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(got["l_hi"], l_ref[:, :, hi * c:(hi + 1) * c], rtol=2e-5, atol=2e-5)


# ---------------------------------------------------------------------------------------------------- ring backward
def test_block_backward_is_an_exact_partial_sum():
    """the rectangle rule the ring backward rests on: with the final L and O, block gradients add up to the full backward"""
    rng = np.random.default_rng(3)
    q, k, v, g = (rng.standard_normal((1, 2, 48, 16)).astype(np.float32) for _ in range(4))
    scale = 0.25
    o, l = O.attention_forward(q, k, v, scale=scale)
    rq, rk, rv, _ = O.attention_backward(q, k, v, g, scale=scale)
    dq = np.zeros_like(q)
    dk, dv = np.zeros_like(k), np.zeros_like(v)
    for k0, k1 in ((0, 20), (20, 48)):
        a, b, c = ring.block_backward_numpy(q, k[:, :, k0:k1], v[:, :, k0:k1], o, l, g, False, scale)
        dq += a
        dk[:, :, k0:k1] += b
        dv[:, :, k0:k1] += c
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5)


def _worker_bwd(rank, world, port, N, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(9)
        B, H, D = 1, 2, 32
        q, k, v, g = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(4))
        scale = 1.0 / np.sqrt(D)
        o, l = O.attention_forward(q, k, v, causal=True, scale=scale)
        be = ring.HostBackend(None, dist)
        sh = lambda x: ring.shard_sequence(np.asarray(x, np.float32), rank, world)
        dq, dk, dv = ring.ring_attention_backward(be, sh(q), sh(k), sh(v), sh(o), sh(l), sh(g), rank, world, scale)
        np.savez(os.path.join(out_dir, f"bwd{rank}.npz"), dq_lo=dq[0], dq_hi=dq[1], dk_lo=dk[0], dk_hi=dk[1], dv_lo=dv[0], dv_hi=dv[1])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 96), (4, 128)])
def test_ring_backward_gloo_matches_single_device(tmp_path, world, N):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker_bwd, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(9)
    B, H, D = 1, 2, 32
    q, k, v, g = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(4))
    rq, rk, rv, _ = O.attention_backward(q, k, v, g, causal=True)
    c = N // (2 * world)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"bwd{r}.npz"))
        lo, hi = ring.chunk_ids(r, world)
        for name, ref in (("dq", rq), ("dk", rk), ("dv", rv)):
            np.testing.assert_allclose(got[f"{name}_lo"], ref[:, :, lo * c:(lo + 1) * c], rtol=3e-4, atol=3e-5)
            np.testing.assert_allclose(got[f"{name}_hi"], ref[:, :, hi * c:(hi + 1) * c], rtol=3e-4, atol=3e-5)
