#!/usr/bin/env python
"""FLUX-shape forward (and forward+backward) under external masks on the tensor-core path: dense additive masks cost mask
bandwidth, structured masks (key padding, sequence packing) skip hidden KV tiles.  Device time from mfa_get_gpu_latency
(CUDA events on the library stream around every kernel of the call, mask pre-pass included); masks are device-resident."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200")); sys.path.insert(0, ROOT)
import torch
import umfa
from umfa import _ffi
lib = _ffi._lib
B, H, S, D = 1, 24, 4608, 128
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
ctx = umfa.MFAContext()
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.randn(B, H, S, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
o = torch.empty(B, H, S, D, device=dev, dtype=torch.float32)
l = torch.empty(B, H, S, device=dev, dtype=torch.float32)
bufs = [umfa.MFABuffer(ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, o, l)]
h = [b.handle for b in bufs]
scale = 1.0 / np.sqrt(D)
i64 = ctypes.c_int64


def fwd(mask, mtype, mscalar):
    if mask is None:
        return lib.mfa_attention_forward_ex(ctx.handle, *h, B, S, S, H, D, scale, False, -1, 1, 2, None, 0, None, None, 0, 0, 0, None)
    shape = (i64 * mask.dim())(*mask.shape)
    strides = (i64 * mask.dim())(*mask.stride())
    return lib.mfa_attention_forward_ex(ctx.handle, *h, B, S, S, H, D, scale, False, -1, 1, 2, ctypes.c_void_p(mask.data_ptr()),
                                        mask.numel() * mask.element_size(), shape, strides, mask.dim(), mtype, mscalar, None)


def visible_fraction(mask):
    if mask is None:
        return 1.0
    m = mask if mask.dtype == torch.bool else torch.isfinite(mask)
    return float(m.expand(B, H, S, S)[0, 0].float().mean()) if m.shape[1] == 1 else float(m.float().mean())


seg = torch.arange(S, device=dev) // (S // 8)
cases = {
    "no_mask": (None, 0, 0),
    "bool_key_padding_[B,1,1,S]_last_512_hidden": (torch.cat([torch.ones(S - 512, dtype=torch.bool, device=dev), torch.zeros(512, dtype=torch.bool, device=dev)]).view(1, 1, 1, S), 1, 0),
    "bool_packing_8_segments_[1,1,S,S]": ((seg[:, None] == seg[None, :]).view(1, 1, S, S), 1, 0),
    "additive_fp32_dense_[1,1,S,S]": (torch.randn(1, 1, S, S, device=dev), 2, 3),
    "additive_bf16_dense_[1,1,S,S]": (torch.randn(1, 1, S, S, device=dev).to(torch.bfloat16), 2, 2),
    "bool_dense_random_half_[1,H,S,S]": (torch.rand(1, H, S, S, device=dev) > 0.5, 1, 0),
    "additive_bf16_dense_[1,H,S,S]": (torch.randn(1, H, S, S, device=dev).to(torch.bfloat16), 2, 2),
    "additive_fp32_dense_[1,H,S,S]": (torch.randn(1, H, S, S, device=dev), 2, 3),
}
res = {}
for name, (mask, mt, ms) in cases.items():
    ts = []
    for i in range(steps + 3):
        rc = fwd(mask, mt, ms)
        assert rc == 0, (name, rc)
        if i >= 3:
            ts.append(ctx.gpu_latency)
    t = float(np.median(ts))
    vis = visible_fraction(mask)
    res[name] = {"ms": t * 1e3, "kernel": ctx.last_kernel, "visible_fraction": vis,
                 "tflops_of_visible_pairs": 4.0 * B * H * S * S * D * vis / t / 1e12,
                 "mask_mb": 0 if mask is None else mask.numel() * mask.element_size() / 1e6}
if os.environ.get("MFA_BENCH_MASK_FWD_ONLY"):
    print(json.dumps({"workload": "FLUX.1-schnell shape B=1 H=24 N=4608 D=128 bf16 forward under external masks",
                      "timing": "mfa_get_gpu_latency, median of %d" % steps, **res}))
    sys.exit(0)
# backward under the same masks (dO random, O / L from the forward just run)
do = torch.randn(B, H, S, D, device=dev, generator=g).to(torch.bfloat16)
dq, dk, dv = (torch.empty(B, H, S, D, device=dev, dtype=torch.float32) for _ in range(3))
dt = torch.empty(B, H, S, device=dev, dtype=torch.float32)
gb = [umfa.MFABuffer(ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (do, dq, dk, dv, dt)]


def bwd(mask, mtype, mscalar):
    margs = [None, 0, None, None, 0, 0, 0]
    if mask is not None:
        margs = [ctypes.c_void_p(mask.data_ptr()), mask.numel() * mask.element_size(), (i64 * mask.dim())(*mask.shape),
                 (i64 * mask.dim())(*mask.stride()), mask.dim(), mtype, mscalar]
    return lib.mfa_attention_backward_ex(ctx.handle, gb[0].handle, h[0], h[1], h[2], h[3], h[4], gb[1].handle, gb[2].handle,
                                         gb[3].handle, gb[4].handle, B, S, S, H, D, scale, False, -1, 1, *margs, None)


for name, (mask, mt, ms) in cases.items():
    assert fwd(mask, mt, ms) == 0
    ts = []
    for i in range(steps + 2):
        rc = bwd(mask, mt, ms)
        assert rc == 0, (name, rc)
        if i >= 2:
            ts.append(ctx.gpu_latency)
    t = float(np.median(ts))
    res[name]["bwd_ms"] = t * 1e3
    res[name]["bwd_kernel"] = ctx.last_kernel
print(json.dumps({"workload": "FLUX.1-schnell shape B=1 H=24 N=4608 D=128 bf16 forward under external masks",
                  "timing": "mfa_get_gpu_latency, median of %d" % steps, **res}))
