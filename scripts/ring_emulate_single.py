#!/usr/bin/env python
"""Compute-only cost of one rank of the single-launch ring (csrc/ring.cu) on ONE GPU: the visiting slots are pre-filled and their
flags pre-raised, so the number is the kernel alone (no exchange), with the same SMs reserved as in the real run.
usage: ring_emulate_single.py [N=131072] [H=32] [worlds=1,8] [reserve=8]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
import umfa
from umfa import ring
from umfa._ffi import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
H = int(sys.argv[2]) if len(sys.argv) > 2 else 32
worlds = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,8").split(",")]
reserve = int(sys.argv[4]) if len(sys.argv) > 4 else 8
D = 128
dev = torch.device("cuda", 0)
ctx = umfa.MFAContext()
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for W in worlds:
    C = N // (2 * W)
    mk = lambda *shape: torch.randn(*shape, device=dev).to(torch.bfloat16)
    q, k, v = mk(1, H, 2 * C, D), mk(1, H, 2 * C, D), mk(1, H, 2 * C, D)
    out = torch.empty(1, H, 2 * C, D, device=dev, dtype=torch.float32)
    lse = torch.empty(1, H, 2 * C, device=dev, dtype=torch.float32)
    kvis = mk(max(W - 1, 1), 1, H, 2 * C, D)
    vvis = mk(max(W - 1, 1), 1, H, 2 * C, D)
    flags = torch.full((W + 1,), 1, device=dev, dtype=torch.int32)
    bufs = [umfa.MFABuffer(ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, out, lse)]
    for rank in sorted({0, W - 1}):
        def run():
            rc = _lib.mfa_attention_forward_ring_slots(ctx.handle, *[b.handle for b in bufs], ctypes.c_void_p(kvis.data_ptr()),
                                                       ctypes.c_void_p(vvis.data_ptr()), ctypes.c_void_p(flags.data_ptr()), 1, rank, W,
                                                       1, C, H, D, 0.088, 1, reserve if W > 1 else 0, st)
            assert rc == 0, rc
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 4.0 * H * ring.visible_pairs_causal(N) * D / W
        print(f"world {W} rank {rank}: {ms:.2f} ms per forward = {flops / ms / 1e9:.0f} TFLOP/s per GPU (single launch, reserve {reserve if W > 1 else 0} SMs)", flush=True)
    del q, k, v, out, lse, kvis, vvis, bufs
    torch.cuda.empty_cache()
