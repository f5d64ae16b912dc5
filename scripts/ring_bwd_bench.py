#!/usr/bin/env python
"""Native ring attention forward + backward at config 5's length (128k causal, H=32, D=128, bf16) under torchrun: CUDA-event time of
the backward (max over ranks) and of forward + backward.  usage: torchrun --nproc-per-node N scripts/ring_bwd_bench.py [N_tokens]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ["MFA_CUDA_DEVICE"] = str(local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import umfa
from umfa import ring
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
B, H, D = 1, 32, 128
dev = torch.device("cuda", local)
C = N // (2 * world)
g = torch.Generator(device=dev).manual_seed(5 + rank)
mk = lambda: tuple(torch.randn(B, H, C, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
ctx = umfa.MFAContext()
os.environ.setdefault("MFA_RING_TRANSPORT", "nccl")
runner = ring.make_runner(ctx, dist if world > 1 else None, dev, "bf16", rank, world)
pk = runner.pack(mk(), mk(), mk())
do = mk()
scale = 1.0 / np.sqrt(D)
def timed(fn, n):
    ts = []
    for i in range(n + 1):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(dev)
        if i > 0: ts.append(a.elapsed_time(b))
    t = torch.tensor([float(np.median(ts))], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
fwd_ms = timed(lambda: runner.forward_packed(pk, scale), 3)
bwd_ms = timed(lambda: runner.backward_packed(pk, do, scale), 3)
pairs = N * (N + 1) / 2
f_fwd, f_bwd = 4.0 * B * H * pairs * D, 10.0 * B * H * pairs * D
if rank == 0:
    print(json.dumps({"workload": f"ring attention {N} tokens causal bf16 H=32 D=128, native driver", "n_gpus": world, "transport": runner.transport,
                      "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "fwd_tflops": f_fwd / fwd_ms / 1e9, "bwd_tflops": f_bwd / bwd_ms / 1e9,
                      "fwdbwd_tflops": (f_fwd + f_bwd) / (fwd_ms + bwd_ms) / 1e9}))
runner.close()
if world > 1: dist.destroy_process_group()
