#!/usr/bin/env python
"""Absolute event list of a few KV steps from a raw MFA_FWD_TRACE file (see attn_fwd_tc.cu launch_traced for the stamps).
usage: fwd_trace_events.py <trace file> [first step] [n steps]"""
import sys
import numpy as np
rows = np.loadtxt(sys.argv[1], dtype=np.float64)
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 2
names = {0: "S ready (softmax)", 1: "S in regs", 2: "turn taken", 3: "P0 pub", 4: "P1 pub", 5: "P2 pub", 6: "P3 pub", 7: "max done (before turn barrier)",
         8: "mma sees P0", 9: "mma sees P1", 10: "mma sees P2", 11: "mma sees P3", 12: "next S issued", 13: "V landed (mma)", 14: "K landed (mma)", 15: "converted (int8)"}
ev = []
for r in rows:
    t, it = int(r[0]), int(r[1])
    if s0 <= it < s0 + ns:
        for k, nm in names.items():
            if r[2 + k] > 0:
                ev.append((r[2 + k], t, it, nm))
ev.sort()
base = ev[0][0]
for c, t, it, nm in ev:
    print(f"{c - base:8.0f}  {'    ' * 6 * t}tile{t} step{it:3d}  {nm}")
