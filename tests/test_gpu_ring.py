"""GPU tests of ring attention (umfa/ring.py CudaBackend over libMFAFFI.so): world=1 on one GPU (3 chunk-pair
kernels + merge) and world=2 over NCCL when two GPUs are visible.  Checked against the CPU oracle's causal attention."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_rank(rank, world, port, N, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    os.environ["MFA_CUDA_DEVICE"] = str(rank)
    torch.cuda.set_device(rank)
    import umfa
    from umfa import ring
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(11)
    B, H, D = 1, 2, 128
    q, k, v = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(3))
    ctx = umfa.MFAContext()
    be = ring.CudaBackend(ctx, dist if world > 1 else None, dev, "bf16")
    sh = lambda x: tuple(torch.from_numpy(np.ascontiguousarray(c)).to(dev).to(torch.bfloat16).contiguous()
                         for c in ring.shard_sequence(x, rank, world))
    (o_lo, l_lo), (o_hi, l_hi) = ring.ring_attention_forward(be, sh(q), sh(k), sh(v), rank, world, 1.0 / np.sqrt(D))
    torch.cuda.synchronize(dev)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), o_lo=o_lo.cpu().numpy(), l_lo=l_lo.cpu().numpy(),
             o_hi=o_hi.cpu().numpy(), l_hi=l_hi.cpu().numpy())
    if world > 1:
        dist.destroy_process_group()


def _check(tmp_path, world, N):
    rng = np.random.default_rng(11)
    B, H, D = 1, 2, 128
    q, k, v = (O.round_bf16(rng.standard_normal((B, H, N, D)).astype(np.float32))[0] for _ in range(3))
    o_ref, l_ref = O.attention_forward(q, k, v, causal=True)
    from umfa import ring
    c = N // (2 * world)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        lo, hi = ring.chunk_ids(r, world)
        for name, cid in (("lo", lo), ("hi", hi)):
            ref = o_ref[:, :, cid * c:(cid + 1) * c]
            err = np.abs(got[f"o_{name}"] - ref).max() / np.abs(ref).max()
            assert err < 2e-2, (r, name, err)
            assert np.abs(got[f"l_{name}"] - l_ref[:, :, cid * c:(cid + 1) * c]).max() < 2e-2


def test_ring_world1(tmp_path):
    _run_rank(0, 1, 0, 1024, str(tmp_path))
    _check(tmp_path, 1, 1024)


def test_ring_world2_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_run_rank, args=(2, port, 2048, str(tmp_path)), nprocs=2, join=True)
    _check(tmp_path, 2, 2048)


def _bf16_dev(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(torch.bfloat16).contiguous()


def test_accumulate_matches_joint_attention():
    """mfa_attention_forward_accumulate: a second K/V block merged in the kernel epilogue into a row window of the running
    (O, L) equals attention over the union of the keys; rows outside the window keep the first block's result."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    import umfa
    from umfa import ring
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    B, H, T, D, K1, K2, W0 = 2, 3, 512, 128, 256, 384, 256
    q = rng.standard_normal((B, H, T, D)).astype(np.float32)
    k = rng.standard_normal((B, H, K1 + K2, D)).astype(np.float32)
    v = rng.standard_normal((B, H, K1 + K2, D)).astype(np.float32)
    ctx = umfa.MFAContext()
    be = ring.CudaBackend(ctx, None, dev, "bf16")
    qd, kd, vd = (_bf16_dev(x, dev) for x in (q, k, v))
    acc = be.attend(qd, kd[:, :, :K1], vd[:, :, :K1], False, 1.0 / np.sqrt(D))
    assert ctx.last_kernel.startswith("fwd_tc_")
    be.attend_accumulate(acc, W0, qd[:, :, W0:], kd[:, :, K1:], vd[:, :, K1:], False, 1.0 / np.sqrt(D))
    torch.cuda.synchronize(dev)
    o, l = acc[0].cpu().numpy(), acc[1].cpu().numpy()
    qr, kr, vr = (O.round_bf16(x)[0] for x in (q, k, v))
    o1, l1 = O.attention_forward(qr[:, :, :W0], kr[:, :, :K1], vr[:, :, :K1])
    o2, l2 = O.attention_forward(qr[:, :, W0:], kr, vr)
    assert np.abs(o[:, :, :W0] - o1).max() / np.abs(o1).max() < 2e-2 and np.abs(l[:, :, :W0] - l1).max() < 2e-2
    assert np.abs(o[:, :, W0:] - o2).max() / np.abs(o2).max() < 2e-2 and np.abs(l[:, :, W0:] - l2).max() < 2e-2
    # the fused merge agrees with the stand-alone merge kernel on the same partials
    acc2 = be.attend(qd, kd[:, :, :K1], vd[:, :, :K1], False, 1.0 / np.sqrt(D))
    part = be.attend(qd[:, :, W0:], kd[:, :, K1:], vd[:, :, K1:], False, 1.0 / np.sqrt(D))
    sub = [acc2[0][:, :, W0:].contiguous(), acc2[1][:, :, W0:].contiguous()]
    be.merge(sub, part)
    torch.cuda.synchronize(dev)
    assert (sub[0] - acc[0][:, :, W0:]).abs().max().item() < 1e-5
    assert (sub[1] - acc[1][:, :, W0:]).abs().max().item() < 1e-5
    ctx.close()


@pytest.mark.parametrize("world", [2, 4])
def test_ring_emulated_on_one_gpu(world):
    """All ranks of a ring run one after the other on one GPU (the exchange hands over the packed K/V of the source rank):
    exercises step_plan, strided row windows and the fused accumulate for every (rank, step) against the oracle."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    import umfa
    from umfa import ring
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(17)
    B, H, D, N = 1, 2, 128, 256 * 2 * world
    q, k, v = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(3))
    o_ref, l_ref = O.attention_forward(*(O.round_bf16(x)[0] for x in (q, k, v)), causal=True)
    ctx = umfa.MFAContext()
    shard = lambda x, r: tuple(_bf16_dev(c, dev) for c in ring.shard_sequence(x, r, world))

    class Emulated(ring.CudaBackend):
        def __init__(self, rank):
            super().__init__(ctx, None, dev, "bf16")
            self.rank = rank
            self.packed = [torch.stack([self.cat_seq(shard(k, r)), self.cat_seq(shard(v, r))]).contiguous()
                           for r in range(world)]

        def exchange_start(self, buf, dst, src, step):
            return self.packed[(self.rank - step - 1) % world]

        def exchange_finish(self, handle, step):
            return handle

    c = N // (2 * world)
    for r in range(world):
        be = Emulated(r)
        (o_lo, l_lo), (o_hi, l_hi) = ring.ring_attention_forward(be, shard(q, r), shard(k, r), shard(v, r), r, world, 1.0 / np.sqrt(D))
        torch.cuda.synchronize(dev)
        lo, hi = ring.chunk_ids(r, world)
        for (o, l), cid in (((o_lo, l_lo), lo), ((o_hi, l_hi), hi)):
            ref = o_ref[:, :, cid * c:(cid + 1) * c]
            assert np.abs(o.cpu().numpy() - ref).max() / np.abs(ref).max() < 2e-2, (r, cid)
            assert np.abs(l.cpu().numpy() - l_ref[:, :, cid * c:(cid + 1) * c]).max() < 2e-2, (r, cid)
    ctx.close()


def test_accumulate_many_items_per_cta():
    """Accumulate mode with more work items than SMs: every persistent CTA merges several items in a row, so the epilogue
    warpgroup's staging ring, the running-O TMA loads and the TMEM hand-back are exercised across item boundaries."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    import umfa
    from umfa import ring
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    B, H, T, D, K1, K2, W0 = 1, 20, 4096 + 60, 128, 256, 320, 1024          # ragged last query block on purpose
    q = rng.standard_normal((B, H, T, D)).astype(np.float32)
    k = rng.standard_normal((B, H, K1 + K2, D)).astype(np.float32)
    v = rng.standard_normal((B, H, K1 + K2, D)).astype(np.float32)
    ctx = umfa.MFAContext()
    be = ring.CudaBackend(ctx, None, dev, "bf16")
    qd, kd, vd = (_bf16_dev(x, dev) for x in (q, k, v))
    acc = be.attend(qd, kd[:, :, :K1], vd[:, :, :K1], False, 1.0 / np.sqrt(D))
    be.attend_accumulate(acc, W0, qd[:, :, W0:], kd[:, :, K1:], vd[:, :, K1:], False, 1.0 / np.sqrt(D))
    be.attend_accumulate(acc, 0, qd[:, :, :W0], kd[:, :, K1:], vd[:, :, K1:], False, 1.0 / np.sqrt(D))
    torch.cuda.synchronize(dev)
    o, l = acc[0].cpu().numpy(), acc[1].cpu().numpy()
    o_ref, l_ref = O.attention_forward(*(O.round_bf16(x)[0] for x in (q, k, v)))
    assert np.isfinite(o).all()
    assert np.abs(o - o_ref).max() / np.abs(o_ref).max() < 2e-2
    assert np.abs(l - l_ref).max() < 2e-2
    ctx.close()


def _run_native_rank(rank, world, port, N, out_dir, transport="nccl"):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    os.environ["MFA_CUDA_DEVICE"] = str(rank)
    torch.cuda.set_device(rank)
    import umfa
    from umfa import ring
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(11)
    B, H, D = 1, 2, 128
    q, k, v = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(3))
    ctx = umfa.MFAContext()
    os.environ["MFA_RING_TRANSPORT"] = transport
    runner = ring.make_runner(ctx, dist, dev, "bf16", rank, world)
    assert runner.kind.startswith("native"), runner.kind
    sh = lambda x: tuple(torch.from_numpy(np.ascontiguousarray(c)).to(dev).to(torch.bfloat16).contiguous()
                         for c in ring.shard_sequence(x, rank, world))
    pk = runner.pack(sh(q), sh(k), sh(v))
    for _ in range(3):                                 # repeated: later forwards reuse the slots (ev_done / consumed flags)
        (o_lo, l_lo), (o_hi, l_hi) = runner.forward_packed(pk, 1.0 / np.sqrt(D))
    torch.cuda.synchronize(dev)
    assert runner.transport.lower().startswith(transport), runner.transport
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), o_lo=o_lo.cpu().numpy(), l_lo=l_lo.cpu().numpy(),
             o_hi=o_hi.cpu().numpy(), l_hi=l_hi.cpu().numpy())
    runner.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_ring_native_world2(tmp_path, transport):
    """csrc/ring.cu end to end over NVLink, both transports: needs two GPUs (skipped on a one-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_run_native_rank, args=(2, port, 2048, str(tmp_path), transport), nprocs=2, join=True)
    _check(tmp_path, 2, 2048)


def test_ring_native_world1(tmp_path):
    """the native driver with a world of one (no transport): a single causal launch through mfa_ring_attention_forward"""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    import umfa
    from umfa import ring
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    B, H, D, N = 1, 2, 128, 1024
    q, k, v = (rng.standard_normal((B, H, N, D)).astype(np.float32) for _ in range(3))
    ctx = umfa.MFAContext()
    runner = ring.make_runner(ctx, None, dev, "bf16", 0, 1)
    assert runner.kind.startswith("native"), runner.kind
    sh = lambda x: tuple(_bf16_dev(c, dev) for c in ring.shard_sequence(x, 0, 1))
    (o_lo, l_lo), (o_hi, l_hi) = runner.forward(sh(q), sh(k), sh(v), 1.0 / np.sqrt(D))
    torch.cuda.synchronize(dev)
    np.savez(os.path.join(str(tmp_path), "rank0.npz"), o_lo=o_lo.cpu().numpy(), l_lo=l_lo.cpu().numpy(),
             o_hi=o_hi.cpu().numpy(), l_hi=l_hi.cpu().numpy())
    runner.close()
    _check(tmp_path, 1, 1024)


# ---------------------------------------------------------------------------------------------------- ring backward
@pytest.mark.parametrize("world", [2, 4])
def test_ring_backward_emulated_on_one_gpu(world):
    """All ranks of a ring as threads on one GPU (the hops are queues between the threads): step_plan rectangles through the
    tensor-core backward kernels with the forward's final O / L, dQ accumulated locally, dK / dV travelling with their K / V
    and arriving home after `world` hops -- against the oracle's full causal backward."""
    import queue
    import threading
    import torch
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    import umfa
    from umfa import ring
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(23)
    B, H, D, N = 1, 2, 128, 256 * 2 * world
    q, k, v, g = (O.round_bf16(rng.standard_normal((B, H, N, D)).astype(np.float32))[0] for _ in range(4))
    o_ref, l_ref = O.attention_forward(q, k, v, causal=True)
    rq, rk, rv, _ = O.attention_backward(q, k, v, g, causal=True)
    ctx = umfa.MFAContext()
    links = [queue.Queue() for _ in range(world)]           # links[r]: what rank r receives from rank r - 1

    class Threaded(ring.CudaBackend):
        def __init__(self, rank):
            super().__init__(ctx, None, dev, "bf16")
            self.rank = rank

        def sendrecv(self, buf, dst, src):
            torch.cuda.synchronize(dev)
            links[dst].put(buf.clone())
            return links[self.rank].get(timeout=120)

    results, errors = {}, []

    def run(r):
        try:
            with torch.cuda.device(dev):
                be = Threaded(r)
                sh16 = lambda x: tuple(_bf16_dev(c, dev) for c in ring.shard_sequence(x, r, world))
                sh32 = lambda x: tuple(torch.from_numpy(np.ascontiguousarray(c, np.float32)).to(dev) for c in ring.shard_sequence(x, r, world))
                results[r] = ring.ring_attention_backward(be, sh16(q), sh16(k), sh16(v), sh32(o_ref), sh32(l_ref), sh16(g), r, world,
                                                          1.0 / np.sqrt(D))
                torch.cuda.synchronize(dev)
        except Exception as e:                                   # noqa: BLE001 -- reported by the main thread
            errors.append((r, repr(e)))

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    c = N // (2 * world)
    for r in range(world):
        lo, hi = ring.chunk_ids(r, world)
        for pair, ref, name in zip(results[r], (rq, rk, rv), ("dq", "dk", "dv")):
            for part, cid in zip(pair, (lo, hi)):
                want = ref[:, :, cid * c:(cid + 1) * c]
                err = np.abs(part.cpu().numpy() - want).max() / np.abs(ref).max()
                assert err < 2e-2, (r, name, cid, err)
    ctx.close()


def _run_rank_bwd(rank, world, port, N, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    os.environ["MFA_CUDA_DEVICE"] = str(rank)
    torch.cuda.set_device(rank)
    import umfa
    from umfa import ring
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(29)
    B, H, D = 1, 2, 128
    q, k, v, g = (O.round_bf16(rng.standard_normal((B, H, N, D)).astype(np.float32))[0] for _ in range(4))
    ctx = umfa.MFAContext()
    be = ring.CudaBackend(ctx, dist, dev, "bf16")
    sh16 = lambda x: tuple(_bf16_dev(c, dev) for c in ring.shard_sequence(x, rank, world))
    (o_lo, l_lo), (o_hi, l_hi) = ring.ring_attention_forward(be, sh16(q), sh16(k), sh16(v), rank, world, 1.0 / np.sqrt(D))
    dq, dk, dv = ring.ring_attention_backward(be, sh16(q), sh16(k), sh16(v), (o_lo, o_hi), (l_lo, l_hi), sh16(g), rank, world,
                                              1.0 / np.sqrt(D))
    torch.cuda.synchronize(dev)
    np.savez(os.path.join(out_dir, f"bwd{rank}.npz"), **{f"{n}_{s}": t.cpu().numpy() for n, pair in (("dq", dq), ("dk", dk), ("dv", dv))
                                                         for s, t in zip(("lo", "hi"), pair)})
    dist.destroy_process_group()


def test_ring_forward_backward_world2_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from umfa import ring
    world, N = 2, 2048
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_run_rank_bwd, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(29)
    q, k, v, g = (O.round_bf16(rng.standard_normal((1, 2, N, 128)).astype(np.float32))[0] for _ in range(4))
    rq, rk, rv, _ = O.attention_backward(q, k, v, g, causal=True)
    c = N // (2 * world)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"bwd{r}.npz"))
        lo, hi = ring.chunk_ids(r, world)
        for name, ref in (("dq", rq), ("dk", rk), ("dv", rv)):
            for s_, cid in (("lo", lo), ("hi", hi)):
                err = np.abs(got[f"{name}_{s_}"] - ref[:, :, cid * c:(cid + 1) * c]).max() / np.abs(ref).max()
                assert err < 2e-2, (r, name, s_, err)


def _run_native_rank_bwd(rank, world, port, N, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    os.environ["MFA_CUDA_DEVICE"] = str(rank)
    torch.cuda.set_device(rank)
    import umfa
    from umfa import ring
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(31)
    q, k, v, g = (O.round_bf16(rng.standard_normal((1, 2, N, 128)).astype(np.float32))[0] for _ in range(4))
    ctx = umfa.MFAContext()
    os.environ["MFA_RING_TRANSPORT"] = "nccl"
    runner = ring.make_runner(ctx, dist if world > 1 else None, dev, "bf16", rank, world)
    assert runner.kind.startswith("native"), runner.kind
    sh = lambda x: tuple(_bf16_dev(c, dev) for c in ring.shard_sequence(x, rank, world))
    pk = runner.pack(sh(q), sh(k), sh(v))
    runner.forward_packed(pk, 1.0 / np.sqrt(128))
    for _ in range(2):                                 # repeated: the second call reuses the work space and the events
        dq, dk, dv = runner.backward_packed(pk, sh(g), 1.0 / np.sqrt(128))
    torch.cuda.synchronize(dev)
    np.savez(os.path.join(out_dir, f"bwd{rank}.npz"), **{f"{n}_{s}": t.cpu().numpy() for n, pair in (("dq", dq), ("dk", dk), ("dv", dv))
                                                         for s, t in zip(("lo", "hi"), pair)})
    runner.close()
    if world > 1:
        dist.destroy_process_group()


def _check_bwd(tmp_path, world, N, seed):
    from umfa import ring
    rng = np.random.default_rng(seed)
    q, k, v, g = (O.round_bf16(rng.standard_normal((1, 2, N, 128)).astype(np.float32))[0] for _ in range(4))
    rq, rk, rv, _ = O.attention_backward(q, k, v, g, causal=True)
    c = N // (2 * world)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"bwd{r}.npz"))
        lo, hi = ring.chunk_ids(r, world)
        for name, ref in (("dq", rq), ("dk", rk), ("dv", rv)):
            for s_, cid in (("lo", lo), ("hi", hi)):
                err = np.abs(got[f"{name}_{s_}"] - ref[:, :, cid * c:(cid + 1) * c]).max() / np.abs(ref).max()
                assert err < 2e-2, (r, name, s_, err)


def test_ring_native_backward_world1(tmp_path):
    """mfa_ring_attention_backward with a world of one: the causal dK/dV + dQ launches through the ring entry point"""
    _run_native_rank_bwd(0, 1, 0, 1024, str(tmp_path))
    _check_bwd(tmp_path, 1, 1024, 31)


def test_ring_native_backward_world2(tmp_path):
    """csrc/ring.cu backward over NVLink (NCCL): K / V exchange, per-step rectangles, dK / dV sent straight back to their owner"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_run_native_rank_bwd, args=(2, port, 2048, str(tmp_path)), nprocs=2, join=True)
    _check_bwd(tmp_path, 2, 2048, 31)


def test_ring_native_backward_world4(tmp_path):
    """four ranks: every kind of step (source below / above the rank, wrap-around) and the reuse of the two send sets"""
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_run_native_rank_bwd, args=(4, port, 4096, str(tmp_path)), nprocs=4, join=True)
    _check_bwd(tmp_path, 4, 4096, 31)
