#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/r02aj_launches_quant.csv python scripts/bench_quant.py 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/r02aj_launches_bwd.csv python bench.py --mode fwdbwd --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras none > /dev/null 2>&1
python - <<PY
import csv, collections
for f in ("$OUT/r02aj_launches_quant.csv", "$OUT/r02aj_launches_bwd.csv"):
    rows=[r for r in csv.reader(l for l in open(f) if not l.startswith("=="))]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    agg=collections.OrderedDict()
    for r in rows[1:]:
        k=r[ki][:64]
        if any(x in k for x in ("quant_span","quant_flat","absmax","codes_to","dterm","head_vscale","fwd_tc","bwd_d")):
            agg.setdefault(k, []).append(float(r[vi].replace(",",""))/1e3)
    for k,v in agg.items(): print(f"{k:64s} n={len(v):3d} median {sorted(v)[len(v)//2]:8.2f} us")
PY
