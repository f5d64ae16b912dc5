#!/bin/bash
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/r02d_gpu_tests.log; cat $OUT/r02d_gpu_tests.log
MFA_DEBUG=1 timeout 120 python scripts/causal_repro.py 12 4608 3 2>&1 | tail -8
MFA_DEBUG=1 timeout 120 python scripts/causal_repro.py 32 16384 2 2>&1 | tail -8
( MFA_DEBUG=1 timeout 300 compute-sanitizer --tool memcheck python scripts/causal_repro.py 32 16384 1 2>&1 | grep -v "^=========     at\|^=========         by\|Host Frame" | head -60 ) > $OUT/r02d_memcheck_causal.txt; head -40 $OUT/r02d_memcheck_causal.txt
export MFA_LIBRARY=$PWD/lib_variants/nostage/libMFAFFI.so
MFA_DEBUG=1 timeout 120 python scripts/causal_repro.py 24 4608 3 2>&1 | tail -8
unset MFA_LIBRARY
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
timeout 200 $B > $OUT/r02d_bench_flux.json 2>>$OUT/r02d_err.txt
timeout 200 $B --workload flux_causal > $OUT/r02d_bench_flux_causal.json 2>>$OUT/r02d_err.txt
timeout 120 python scripts/cta_trace.py flux $OUT/r02d_cta_trace.txt > /dev/null 2>>$OUT/r02d_err.txt
( timeout 120 python scripts/fwd_trace.py bf16 2 2>&1 | tail -8 ) > $OUT/r02d_fwd_trace.txt
python - <<PY
import json
for f in ("flux", "flux_causal"):
    try:
        d=json.load(open("$OUT/r02d_bench_%s.json" % f))
        print(f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
grep -E "loop|epilogue|first_S|kernel span" $OUT/r02d_cta_trace.txt; cat $OUT/r02d_fwd_trace.txt; tail -3 $OUT/r02d_err.txt
