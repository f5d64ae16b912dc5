#!/usr/bin/env python
"""HBM-bound helper kernels on FLUX-shape tensors (B=1 H=24 N=4608 D=128), one call each per round: meant to run under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` (scripts/gpu_r2_call52.sh), which gives the
per-launch time and DRAM bytes; this script itself only issues the calls.  Helpers: Hadamard rotation (warp kernel and the
shared-memory kernel), the D-term pre-pass of the backward, block-64 / per-tensor int8 quantisers (through the runtime-quantised
forward)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200")); sys.path.insert(0, ROOT)
import torch
import umfa
from umfa import _ffi
lib = _ffi._lib
B, H, S, D = 1, 24, 4608, 128
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
ctx = umfa.MFAContext()
g = torch.Generator(device=dev).manual_seed(0)
x32 = torch.randn(B, H, S, D, device=dev, generator=g)
hb = umfa.MFABuffer(ctx, device_ptr=x32.data_ptr(), size=x32.numel() * 4)
q, k, v = (torch.randn(B, H, S, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
o = torch.empty(B, H, S, D, device=dev, dtype=torch.float32)
l = torch.empty(B, H, S, device=dev, dtype=torch.float32)
bufs = [umfa.MFABuffer(ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, o, l)]
h = [b.handle for b in bufs]
scale = 1.0 / np.sqrt(D)
for r in range(rounds):
    for n in (128, 64, 1024):
        assert lib.mfa_hadamard_rotate(hb.handle, n, x32.numel() // n) == 0
    os.environ["MFA_HADAMARD_SMEM"] = "1"
    assert lib.mfa_hadamard_rotate(hb.handle, 128, x32.numel() // 128) == 0
    del os.environ["MFA_HADAMARD_SMEM"]
    for mode in (0, 2):          # per-tensor, block-64 int8 codes
        assert lib.mfa_quantized_forward_with_lse(ctx.handle, *h, None, B, S, S, H, D, scale, False, 3, mode, 1) == 0
torch.cuda.synchronize()
print("ok")
