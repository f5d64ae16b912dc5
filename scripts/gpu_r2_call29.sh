#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --extras none"
for i in 1 2 3; do
  for v in r02r head; do
    if [ $v = head ]; then L=""; else L="MFA_LIBRARY=$PWD/lib_variants/$v/libMFAFFI.so"; fi
    env $L timeout 200 python $B > $OUT/r02ad_${v}_${i}.json 2>>$OUT/r02ad_err.txt
    python - <<PY
import json
d=json.loads(open("$OUT/r02ad_${v}_${i}.json").read().strip().splitlines()[-1])
print("$v", $i, round(d["value"],1), round(d["ms_per_step"],4), d["clocks"]["sm_mhz"])
PY
  done
done
( timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fp32_tc.py tests/test_gpu_ring.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -12 ) > $OUT/r02ad_tests.log; cat $OUT/r02ad_tests.log
tail -2 $OUT/r02ad_err.txt
