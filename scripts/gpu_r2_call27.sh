#!/bin/bash
# round 2: same-box A/B of the bf16 forward: library of commit fa3f8bb (r02r, 1212 TFLOP/s) vs HEAD
OUT=gpurun_out; mkdir -p $OUT
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --extras none"
for i in 1 2 3; do
  MFA_LIBRARY=$PWD/lib_variants/r02r/libMFAFFI.so timeout 200 python $B > $OUT/r02ab_flux_r02r_$i.json 2>>$OUT/r02ab_err.txt
  timeout 200 python $B > $OUT/r02ab_flux_head_$i.json 2>>$OUT/r02ab_err.txt
done
python - <<PY
import json
for tag in ("r02r", "head"):
    for i in (1,2,3):
        try:
            d=json.loads(open("$OUT/r02ab_flux_%s_%d.json" % (tag,i)).read().strip().splitlines()[-1])
            print(tag, i, round(d["value"],1), round(d["ms_per_step"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        except Exception as e: print(tag, i, "failed", e)
PY
tail -2 $OUT/r02ab_err.txt
