#!/bin/bash
# single-trip block quantiser + two-phase mask pre-pass: quantiser / quantised-attention / mask tests, helper kernels under ncu (single
# trip and two-trip), mask bench + its launch list, quantised bench
TAG=${1:-r02bh}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_quant.py tests/test_gpu_tcq.py tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py -m gpu -q -x 2>&1 | tail -4 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
for M in single two; do
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
print("$M:", {k[-40:]: sorted(v)[len(v)//2] for k, v in agg.items() if "quant" in k or "e4m3" in k or "vscale" in k})
PY
done
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2>> $OUT/${TAG}_err.txt
MFA_BENCH_MASK_FWD_ONLY=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_mask.csv \
    python scripts/bench_mask.py 2 > /dev/null 2>&1
grep "mask_flags" $OUT/${TAG}_launches_mask.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_mask.json").read().strip().splitlines()[-1])
    print({k: (round(v["ms"], 4), round(v.get("bwd_ms", 0), 4)) for k, v in d.items() if isinstance(v, dict)})
except Exception as e: print("mask failed", e)
PY
timeout 300 python scripts/bench_quant.py 10 > $OUT/${TAG}_bench_quant.json 2>> $OUT/${TAG}_err.txt
head -c 1500 $OUT/${TAG}_bench_quant.json; echo
tail -3 $OUT/${TAG}_err.txt
