#!/bin/bash
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/r02i_gpu_tests.log; cat $OUT/r02i_gpu_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
for V in default release; do
  if [ $V = default ]; then unset MFA_LIBRARY; else export MFA_LIBRARY=$PWD/lib_variants/$V/libMFAFFI.so; fi
  for W in flux flux_causal long_dense; do timeout 200 $B --workload $W > $OUT/r02i_bench_${W}_$V.json 2>>$OUT/r02i_err.txt; done
  timeout 120 python scripts/cta_trace.py flux $OUT/r02i_cta_trace_$V.txt > /dev/null 2>>$OUT/r02i_err.txt
  python - <<PY
import json
for f in ("flux", "flux_causal", "long_dense"):
    try:
        d=json.load(open("$OUT/r02i_bench_%s_$V.json" % f))
        print("$V", f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"])
    except Exception as e: print("$V", f, "failed", e)
PY
  grep -E "loop|epilogue|first_S|pv_tail|kernel span" $OUT/r02i_cta_trace_$V.txt
  timeout 300 python scripts/ring_emulate_single.py 131072 32 1,8 8 > $OUT/r02i_ring_single_$V.txt 2>>$OUT/r02i_err.txt; cat $OUT/r02i_ring_single_$V.txt
done
grep "mfa\]" $OUT/r02i_err.txt | head -20; tail -5 $OUT/r02i_err.txt
