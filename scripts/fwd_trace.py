#!/usr/bin/env python
"""Timeline of one CTA of the tcgen05 forward (MFA_FWD_TRACE build): average clocks per phase, steps 5..30.
usage: fwd_trace.py [bf16|int8] [poly]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200")); sys.path.insert(0, ROOT)
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
poly = sys.argv[2] if len(sys.argv) > 2 else "0"
qmode = int(sys.argv[3]) if len(sys.argv) > 3 else 2          # 0 per-tensor scales, 2 per-block (64 tokens)
os.environ["MFA_FWD_POLY"] = poly
import torch
import umfa
from umfa import _ffi
lib = _ffi._lib
B, H, S, D = 1, 24, 4608, 128
dev = torch.device("cuda", 0)
ctx = umfa.MFAContext()
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.randn(B, H, S, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
o = torch.empty(B, H, S, D, device=dev, dtype=torch.float32)
l = torch.empty(B, H, S, device=dev, dtype=torch.float32)
bufs = [umfa.MFABuffer(ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, o, l)]
h = [b.handle for b in bufs]
scale = 1.0 / np.sqrt(D)
def call():
    if mode == "bf16":
        return lib.mfa_attention_forward_with_lse(ctx.handle, *h, B, S, S, H, D, scale, False, 1, 2, False, False, False, False)
    return lib.mfa_quantized_forward_with_lse(ctx.handle, *h, None, B, S, S, H, D, scale, False, 3, qmode, 1)
for _ in range(3):
    assert call() == 0
path = os.path.join(ROOT, "gpurun_out", f"fwd_trace_{mode}_{poly}_q{qmode}.txt"); os.makedirs(os.path.dirname(path), exist_ok=True)
os.environ["MFA_FWD_TRACE"] = path
assert call() == 0
del os.environ["MFA_FWD_TRACE"]
rows = np.loadtxt(path, dtype=np.float64)
print(f"== {mode} POLY={poly}: {len(rows)} rows")
for t in (0, 1):
    r = rows[(rows[:, 0] == t) & (rows[:, 1] >= 5) & (rows[:, 1] <= 30)]
    st = r[:, 2:]
    per = np.diff(st[:, 0]).mean()
    names = ["wait_S(prev end->S ready)", "ld", "mask+max+rescale", "part0 pub", "part1 pub", "part2 pub", "part3 pub"]
    d = [np.nan, (st[:, 1] - st[:, 0]).mean(), (st[:, 2] - st[:, 1]).mean()] + [(st[:, 3 + i] - st[:, 2 + i]).mean() for i in range(4)]
    d[0] = (st[1:, 0] - st[:-1, 6]).mean()
    print(f" tile {t}: period {per:.0f} clk | " + " | ".join(f"{n} {x:.0f}" for n, x in zip(names, d)))
    m = [(st[:, 8 + i] - st[:, 3 + i]).mean() for i in range(4)]
    print(f"   MMA: part pub->seen {[round(x) for x in m]} | last part seen -> next S issued+committed {(st[:, 12] - st[:, 11]).mean():.0f}"
          f" | S issue -> softmax sees S {(st[1:, 0] - st[:-1, 12]).mean():.0f} | V wait->part0 seen {(st[:, 8] - st[:, 13]).mean():.0f}")
r0 = rows[(rows[:, 0] == 0) & (rows[:, 1] >= 5) & (rows[:, 1] <= 30)][:, 2:]
r1 = rows[(rows[:, 0] == 1) & (rows[:, 1] >= 5) & (rows[:, 1] <= 30)][:, 2:]
print(f" tile1 S-ready minus tile0 S-ready: {(r1[:, 0] - r0[:, 0]).mean():.0f} clk")
