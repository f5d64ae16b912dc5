#!/bin/bash
# bisect the bf16 forward regression on one box
OUT=gpurun_out; mkdir -p $OUT
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --extras none"
for i in 1 2; do
  for v in r02r c_4febb16 c_7a28c69 head; do
    if [ $v = head ]; then L=""; else L="MFA_LIBRARY=$PWD/lib_variants/$v/libMFAFFI.so"; fi
    env $L timeout 200 python $B > $OUT/r02ac_${v}_${i}.json 2>>$OUT/r02ac_err.txt
    python - <<PY
import json
d=json.loads(open("$OUT/r02ac_${v}_${i}.json").read().strip().splitlines()[-1])
print("$v", $i, round(d["value"],1), round(d["ms_per_step"],4), d["clocks"]["sm_mhz"])
PY
  done
done
tail -2 $OUT/r02ac_err.txt
