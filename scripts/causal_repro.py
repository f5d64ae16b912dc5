#!/usr/bin/env python
"""Causal forward with more items than SMs (persistent CTAs chew through items of very different length) checked against a
torch fp32 reference; prints the library's CUDA error text when MFA_DEBUG=1.  usage: causal_repro.py [H] [S] [iters]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
import numpy as np, torch
import umfa
from umfa import _ffi
H = int(sys.argv[1]) if len(sys.argv) > 1 else 12
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4608
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
D = 128
lib = _ffi._lib
ctx = umfa.MFAContext()
dev = torch.device("cuda", 0)
q, k, v = (torch.randn(1, H, S, D, device=dev).to(torch.bfloat16) for _ in range(3))
o = torch.zeros(1, H, S, D, device=dev, dtype=torch.float32)
l = torch.zeros(1, H, S, device=dev, dtype=torch.float32)
bufs = [umfa.MFABuffer(ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, o, l)]
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for i in range(iters):
    rc = lib.mfa_attention_forward_ex(ctx.handle, *[b.handle for b in bufs], 1, S, S, H, D, 1.0 / np.sqrt(D), True, -1, 1, 2,
                                      None, 0, None, None, 0, 0, 0, st)
    print("launch", i, "rc", rc, flush=True)
    torch.cuda.synchronize()
ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float(), is_causal=True)
print("kernel", ctx.last_kernel, "max rel err", float((o - ref).abs().max() / ref.abs().max()))
