#!/bin/bash
# 8-wide code rule of the quantisers (merged rare branch, no clamps): bit-exact quantiser tests, quantised attention tests, helper launch
# list (single-trip variant and default), quantised bench
TAG=${1:-r02bm}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_quant.py tests/test_gpu_tcq.py -m gpu -q -x 2>&1 | tail -4 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
for M in two single; do
  [ $M = single ] && export MFA_QUANT_SINGLE_TRIP=1
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
      --log-file $OUT/${TAG}_helpers_$M.csv python scripts/bench_helpers.py 3 > /dev/null 2>$OUT/${TAG}_err.txt
  unset MFA_QUANT_SINGLE_TRIP
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/${TAG}_helpers_$M.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][-56:]
    if r[-3] != "gpu__time_duration.sum" or "at::" in name or "distribution" in name: continue
    agg.setdefault(name, []).append(float(r[-1].replace(",", "")))
print("$M:", {k[-40:]: sorted(v)[len(v)//2] for k, v in agg.items() if "quant" in k or "e4m3" in k or "vscale" in k or "absmax" in k})
PY
done
timeout 300 python scripts/bench_quant.py 10 > $OUT/${TAG}_bench_quant.json 2>> $OUT/${TAG}_err.txt
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_quant.json").read().strip().splitlines()[-1])
print({k: (round(v["ms"], 4), round(v.get("speedup_vs_bf16_incl_quantise", 1), 3)) for k, v in d.items() if isinstance(v, dict)})
PY
tail -3 $OUT/${TAG}_err.txt
