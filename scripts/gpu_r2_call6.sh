#!/bin/bash
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
for V in default release; do
  if [ $V = default ]; then unset MFA_LIBRARY; else export MFA_LIBRARY=$PWD/lib_variants/$V/libMFAFFI.so; fi
  for W in flux flux_causal; do timeout 200 $B --workload $W > $OUT/r02g_bench_${W}_$V.json 2>>$OUT/r02g_err.txt; done
  MFA_FWD_PERSIST=0 timeout 200 $B > $OUT/r02g_bench_flux_nopersist_$V.json 2>>$OUT/r02g_err.txt
  MFA_FWD_PINGPONG=0 timeout 200 $B > $OUT/r02g_bench_flux_nopingpong_$V.json 2>>$OUT/r02g_err.txt
  timeout 120 python scripts/cta_trace.py flux $OUT/r02g_cta_trace_$V.txt > /dev/null 2>>$OUT/r02g_err.txt
  ( timeout 120 python scripts/fwd_trace.py bf16 2 2>&1 | tail -8 ) > $OUT/r02g_fwd_trace_$V.txt
  python - <<PY
import json
for f in ("flux", "flux_causal", "flux_nopersist", "flux_nopingpong"):
    try:
        d=json.load(open("$OUT/r02g_bench_%s_$V.json" % f))
        print("$V", f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"])
    except Exception as e: print("$V", f, "failed", e)
PY
  grep -E "loop|epilogue|first_S|pv_tail|kernel span" $OUT/r02g_cta_trace_$V.txt; cat $OUT/r02g_fwd_trace_$V.txt
done
unset MFA_LIBRARY
timeout 200 python scripts/ring_emulate.py 131072 32 8 > $OUT/r02g_ring_emulate.txt 2>>$OUT/r02g_err.txt; cat $OUT/r02g_ring_emulate.txt
tail -3 $OUT/r02g_err.txt
