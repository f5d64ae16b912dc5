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
