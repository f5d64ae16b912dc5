#!/usr/bin/env python
"""Stall samples of one kernel in an .ncu-rep bucketed by SASS region, with the marker instructions of each bucket.
usage: ncu_regions.py report.ncu-rep [bucket]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Kernel Name": hdr = None; data = []; continue
    if hdr is None and r and "Source" in r: hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
idx = {h: i for i, h in enumerate(hdr)}
S = idx["# Samples"]; SRC = idx["Source"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in data)
print("total samples", tot, "lines", len(data))
MARK = re.compile(r"LDTM|STTM|SYNCS|BAR\.|UTC|MUFU\.EX2|I2FP|UTMALDG|STG|LDG|WARPSYNC|EXIT|FMNMX3|F2FP")
for b0 in range(0, len(data), bucket):
    chunk = data[b0:b0 + bucket]
    n = sum(int(r[S]) for r in chunk)
    if n < tot * 0.004: continue
    st = collections.Counter()
    for r in chunk:
        for k in stall_cols:
            v = int(r[idx[k]])
            if v: st[k[6:]] += v
    ops = collections.Counter(m.group(0) for r in chunk for m in [MARK.search(r[SRC])] if m)
    print(f"{b0:5d}-{b0+len(chunk):5d} {n:6d} {100*n/tot:5.1f}%  {dict(st.most_common(4))}  {dict(ops)}")
