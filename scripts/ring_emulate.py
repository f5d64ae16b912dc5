#!/usr/bin/env python
"""Compute-only cost of one rank of the ring (no communication): runs rank r of world W of the config-5 problem on ONE
GPU with the exchange replaced by a pointer swap, timing every ring step with CUDA events.  Separates kernel efficiency at
the ring's sub-problem sizes from communication / synchronisation losses seen in the multi-GPU bench.
usage: ring_emulate.py [N=131072] [H=32] [worlds=1,4,8]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
import umfa
from umfa import ring
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
H = int(sys.argv[2]) if len(sys.argv) > 2 else 32
worlds = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,4,8").split(",")]
D = 128
dev = torch.device("cuda", 0)
ctx = umfa.MFAContext()


class Emulated(ring.CudaBackend):
    def __init__(self):
        super().__init__(ctx, None, dev, "bf16")
        self.ev = []

    def exchange_start(self, buf, dst, src, step):
        return self.kv_bufs[self.cur ^ 1]

    def exchange_finish(self, handle, step):
        e = torch.cuda.Event(enable_timing=True); e.record(self.compute); self.ev.append(e)
        self.cur ^= 1
        return handle


for W in worlds:
    C = N // (2 * W)
    mk = lambda: torch.randn(1, H, C, D, device=dev).to(torch.bfloat16)
    qp, kp, vp = (mk(), mk()), (mk(), mk()), (mk(), mk())
    for rank in sorted({0, W - 1}):
        be = Emulated()
        for _ in range(2):
            ring.ring_attention_forward(be, qp, kp, vp, rank, W, 0.088)
        torch.cuda.synchronize()
        be.kv_bufs[1].normal_()
        reps, tot, steps = 3, 0.0, None
        for _ in range(reps):
            be.ev = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(be.compute)
            ring.ring_attention_forward(be, qp, kp, vp, rank, W, 0.088)
            e1.record(be.compute)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
            marks = [e0] + be.ev + [e1]
            steps = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
        ms = tot / reps
        flops = 4.0 * H * ring.visible_pairs_causal(N) * D / W
        print(f"world {W} rank {rank}: {ms:.2f} ms per forward = {flops / ms / 1e9:.0f} TFLOP/s per GPU; per step ms: "
              + " ".join(f"{x:.2f}" for x in steps), flush=True)
        del be
    del qp, kp, vp
    torch.cuda.empty_cache()
