#!/usr/bin/env python
"""Config 3 of BASELINE.json: FLUX shape, bf16 forward vs the int8 / int4 quantised forward (device time from
mfa_get_gpu_latency: CUDA events on the library stream around all kernels of the call, i.e. INCLUDING the runtime
quantise / unpack / V-convert pre-passes; the ncu launch list in profiles/ splits them)."""
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
flops = 4.0 * B * H * S * S * D
res = {}
def run(name, fn):
    ts = []
    for i in range(steps + 3):
        rc = fn()
        assert rc == 0, (name, rc)
        if i >= 3: ts.append(ctx.gpu_latency)
    t = float(np.median(ts))
    res[name] = {"ms": t * 1e3, "tflops": flops / t / 1e12, "kernel": ctx.last_kernel}
run("bf16", lambda: lib.mfa_attention_forward_with_lse(ctx.handle, *h, B, S, S, H, D, scale, False, 1, 2, False, False, False, False))
ref = o.clone()
for name, tp, mode in (("int8_tensor", 3, 0), ("int8_block64", 3, 2), ("int4_block64", 4, 2)):
    run(name, lambda: lib.mfa_quantized_forward_with_lse(ctx.handle, *h, None, B, S, S, H, D, scale, False, tp, mode, 1))
    a, b_ = o.double().flatten(), ref.double().flatten()
    res[name]["cosine_vs_bf16"] = float((a @ b_) / (a.norm() * b_.norm()))
    res[name]["max_abs_vs_bf16"] = float((a - b_).abs().max())
    res[name]["speedup_vs_bf16_incl_quantise"] = res["bf16"]["ms"] / res[name]["ms"]
print(json.dumps({"workload": "FLUX.1-schnell shape B=1 H=24 N=4608 D=128, forward", "timing": "mfa_get_gpu_latency, median of %d" % steps, **res}))
