"""Generates tests/golden/sdpa_golden.npz with the reference's own Python test oracle.

The reference stores no attention output tensors; its Python adapter tests define correctness as
agreement with torch's CPU scaled_dot_product_attention
(/root/reference/examples/pytorch-custom-op-ffi/tests/conftest.py:165-181 `reference_attention`,
inputs per conftest.py:147-160: torch.manual_seed(42); randn * 0.1; tolerances conftest.py:189-198).
This script restates that recipe, runs it in THIS container (torch CPU) and commits inputs + outputs so
the oracle (and through it the CUDA path) is pinned without /root/reference or torch-version drift at
test time.  Layout is BHSD, the layout the adapters pass (metal_sdpa_backend.cpp:2697-2702).

Run:  python tests/golden/make_golden.py
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sdpa_golden.npz")


def make_inputs(shape_q, shape_kv, dtype=torch.float32, seed=42):
    torch.manual_seed(seed)
    q = torch.randn(shape_q, dtype=torch.float32) * 0.1
    k = torch.randn(shape_kv, dtype=torch.float32) * 0.1
    v = torch.randn(shape_kv, dtype=torch.float32) * 0.1
    # round to the dtype under test, keep fp32 containers (the oracle consumes fp32)
    return [t.to(dtype).to(torch.float32) for t in (q, k, v)]


def main():
    out = {}
    cases = [
        # name, (B,H,Sq,D), Skv, dtype, causal, mask kind, scale, want_grad
        ("c1_fp32", (1, 1, 512, 64), 512, torch.float32, False, None, None, False),
        ("causal_fp32", (2, 3, 96, 32), 96, torch.float32, True, None, None, True),
        ("rect_fp32", (1, 2, 80, 48), 112, torch.float32, False, None, 0.25, True),
        ("boolmask_fp32", (1, 2, 80, 48), 112, torch.float32, False, "bool", None, False),
        ("addmask_fp32", (2, 2, 40, 16), 56, torch.float32, False, "add", None, True),
        ("bf16_d128", (1, 2, 160, 128), 160, torch.bfloat16, False, None, None, False),
        ("fp16_causal_d64", (1, 2, 130, 64), 130, torch.float16, True, None, None, False),
        # sliding window (kernel-source rule, AttentionKernel+Softmax.swift:445,450: hidden iff col > row or row > col + W),
        # handed to torch as the equivalent banded bool mask: the independent cross-check of the window numerics that
        # upstream's compile-only tests do not give.  meta[2] = W.
        ("window_fp32", (1, 2, 200, 32), 200, torch.float32, True, ("window", 37), None, True),
        ("window_bf16_d128", (1, 2, 300, 128), 300, torch.bfloat16, True, ("window", 130), None, False),
    ]
    for name, (B, H, Sq, D), Skv, dtype, causal, mkind, scale, want_grad in cases:
        q, k, v = make_inputs((B, H, Sq, D), (B, H, Skv, D), dtype)
        mask = None
        window = -1
        if isinstance(mkind, tuple):
            window = mkind[1]
            r, c = torch.arange(Sq)[:, None], torch.arange(Skv)[None, :]
            mask = ((c <= r) & ~(r > c + window))[None, None]
            causal = False                      # torch gets the whole rule as a mask; the oracle / kernel get (causal, window)
        if mkind == "bool":
            g = torch.Generator().manual_seed(7)
            mask = torch.rand((1, 1, Sq, Skv), generator=g) > 0.3
            mask[..., 0] = True  # keep every row attendable (torch gives NaN for empty rows)
        elif mkind == "add":
            g = torch.Generator().manual_seed(8)
            mask = torch.randn((B, H, Sq, Skv), generator=g)
        if want_grad:
            q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=causal, scale=scale)
        out[f"{name}.q"] = q.detach().numpy()
        out[f"{name}.k"] = k.detach().numpy()
        out[f"{name}.v"] = v.detach().numpy()
        out[f"{name}.o"] = o.detach().numpy()
        if window >= 0:
            out[f"{name}.meta"] = np.array([1, -1.0 if scale is None else scale, window], np.float64)
        else:
            out[f"{name}.meta"] = np.array([int(causal), -1.0 if scale is None else scale], np.float64)
            if mask is not None:
                out[f"{name}.mask"] = mask.numpy()
        if want_grad:
            torch.manual_seed(43)
            d_o = torch.randn_like(o) * 0.1
            o.backward(d_o)
            out[f"{name}.do"] = d_o.numpy()
            out[f"{name}.dq"] = q.grad.numpy()
            out[f"{name}.dk"] = k.grad.numpy()
            out[f"{name}.dv"] = v.grad.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "torch", torch.__version__)


if __name__ == "__main__":
    main()
