#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics per captured launch and the top stall sites from the
source page.  usage: ncu_summary.py report.ncu-rep [n_top]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "sm__cycles_active.avg", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} [{units[i]}]:", [d[i][:60] for d in data])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if b["name"] in seen:
        continue
    seen.add(b["name"])
    idx = {h: i for i, h in enumerate(b["hdr"])}
    S = idx["# Samples"]
    tot = sum(int(r[S]) for r in b["rows"])
    print(f"\n=== {b['name'][:90]}  samples={tot} sass_lines={len(b['rows'])}")
    stall_cols = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    top = sorted(enumerate(b["rows"]), key=lambda t: -int(t[1][S]))[:ntop]
    for i, r in sorted(top):
        st = {k[6:]: int(r[idx[k]]) for k in stall_cols if int(r[idx[k]]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{i:5d} {r[idx['Source']].strip()[:64]:64s} {r[S]:>6s} {st}")
