#!/usr/bin/env python
"""Per-CTA wall-clock timeline of the tensor-core forward (MFA_FWD_CTATRACE debug build of the launch).

    python scripts/cta_trace.py [workload] [out.txt]

Launches the workload a few times normally, then once with the instrumented kernel, and prints where a CTA's life goes:
set-up, wait for the first S, main loop, last P V, epilogue, exit, and the gap until the next CTA starts on the same SM.
"""
import ctypes
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))

import numpy as np
import torch

import bench
import umfa
from umfa import _ffi


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "flux"
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/cta_trace.txt"
    w = bench.WORKLOADS[wl]
    lib = _ffi._lib
    ctx = umfa.MFAContext()
    dev = torch.device("cuda", 0)
    B, H, Sq, Skv, D = w["B"], w["H"], w["Sq"], w["Skv"], w["D"]
    q, k, v = (torch.randn(B, H, S, D, device=dev).to(torch.bfloat16) for S in (Sq, Skv, Skv))
    o = torch.empty(B, H, Sq, D, device=dev, dtype=torch.float32)
    bufs = [umfa.MFABuffer(ctx, device_ptr=t.data_ptr(), size=t.numel() * t.element_size()) for t in (q, k, v, o)]
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def launch():
        rc = lib.mfa_attention_forward_ex(ctx.handle, bufs[0].handle, bufs[1].handle, bufs[2].handle, bufs[3].handle, None,
                                          B, Sq, Skv, H, D, 1.0 / np.sqrt(D), w["causal"], w["window"], 1, 2,
                                          None, 0, None, None, 0, 0, 0, st)
        assert rc == 0, rc

    for _ in range(10):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        launch()
    e1.record()
    torch.cuda.synchronize()
    print(f"{wl}: plain launch {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
    raw = out + ".raw"
    os.environ["MFA_FWD_CTATRACE"] = raw
    launch()
    torch.cuda.synchronize()
    del os.environ["MFA_FWD_CTATRACE"]
    rows = np.loadtxt(raw, dtype=np.int64)
    t00 = rows[:, 1].min()
    by_sm = defaultdict(list)
    for r in rows:
        by_sm[int(r[7])].append(r)
    names = ["setup", "first_S", "loop", "pv_tail", "epilogue", "exit"]
    seg = {n: [] for n in names}
    gaps, order_stats = [], defaultdict(list)
    for sm, lst in by_sm.items():
        lst.sort(key=lambda r: r[1])
        for i, r in enumerate(lst):
            st_ = [r[1], r[2], r[3], r[4], r[5], r[6], r[10]]
            for n, a, b in zip(names, st_[:-1], st_[1:]):
                seg[n].append(b - a)
            order_stats[i].append((r[1] - t00, r[10] - t00))
            if i + 1 < len(lst):
                gaps.append(lst[i + 1][1] - r[10])
    with open(out, "w") as f:
        def pr(s):
            print(s)
            f.write(s + "\n")
        pr(f"# {wl}: {len(rows)} CTAs on {len(by_sm)} SMs; kernel span {(rows[:, 10].max() - t00) / 1e3:.1f} us (globaltimer)")
        for n in names:
            a = np.array(seg[n]) / 1e3
            pr(f"{n:10s} median {np.median(a):8.2f} us   p10 {np.percentile(a, 10):8.2f}   p90 {np.percentile(a, 90):8.2f}")
        if gaps:
            g = np.array(gaps) / 1e3
            pr(f"{'next_gap':10s} median {np.median(g):8.2f} us   p10 {np.percentile(g, 10):8.2f}   p90 {np.percentile(g, 90):8.2f}")
        for i in sorted(order_stats):
            a = np.array(order_stats[i]) / 1e3
            pr(f"CTA #{i} on its SM: start median {np.median(a[:, 0]):8.2f} us (max {a[:, 0].max():8.2f}), end median {np.median(a[:, 1]):8.2f} us (max {a[:, 1].max():8.2f}), n={len(a)}")
        clk = (rows[:, 9] - rows[:, 8]) / np.maximum(1, rows[:, 6] - rows[:, 1])
        pr(f"SM clock during CTAs (clock64 / globaltimer): median {np.median(clk):.3f} GHz  min {clk.min():.3f}  max {clk.max():.3f}")
    os.remove(raw)


if __name__ == "__main__":
    main()
