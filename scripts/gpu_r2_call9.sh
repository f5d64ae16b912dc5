#!/bin/bash
# same-box A/B: round-1 forward kernel (hybrid library) vs the current one (release build), two passes
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
for PASS in 1 2; do
for V in r1 release; do
  export MFA_LIBRARY=$PWD/lib_variants/$V/libMFAFFI.so
  for W in flux long_dense flux_causal; do timeout 200 $B --workload $W > $OUT/r02j_bench_${W}_${V}_$PASS.json 2>>$OUT/r02j_err.txt; done
  python - <<PY
import json
for f in ("flux", "long_dense", "flux_causal"):
    try:
        d=json.load(open("$OUT/r02j_bench_%s_${V}_$PASS.json" % f))
        print("$V pass $PASS", f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"]["sm_mhz"])
    except Exception as e: print("$V", f, "failed", e)
PY
done
done
for V in r1 release; do
  export MFA_LIBRARY=$PWD/lib_variants/$V/libMFAFFI.so
  ( timeout 120 python scripts/fwd_trace.py bf16 2 2>&1 | tail -8 ) > $OUT/r02j_fwd_trace_$V.txt; echo $V; cat $OUT/r02j_fwd_trace_$V.txt
done
tail -3 $OUT/r02j_err.txt
