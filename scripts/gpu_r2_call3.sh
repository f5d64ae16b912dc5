#!/bin/bash
# round 2, GPU visit 3: accumulate-mode fault hunt (compute-sanitizer) + forward variants (staging / ring depth / register split)
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ring.py -m gpu -x -q -k "many_items" 2>&1 | tail -40 ) > $OUT/r02c_memcheck_acc.txt; tail -25 $OUT/r02c_memcheck_acc.txt
( timeout 300 python -m pytest tests/test_gpu_ring.py tests/test_gpu_tcq.py -m gpu -x -q 2>&1 | tail -8 ) > $OUT/r02c_tests.txt; cat $OUT/r02c_tests.txt
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
for V in default nostage regs216 nostage216; do
  if [ $V = default ]; then unset MFA_LIBRARY; else export MFA_LIBRARY=$PWD/lib_variants/$V/libMFAFFI.so; fi
  timeout 200 $B > $OUT/r02c_bench_flux_$V.json 2>>$OUT/r02c_err.txt
  timeout 200 $B --workload flux_causal > $OUT/r02c_bench_flux_causal_$V.json 2>>$OUT/r02c_err.txt
  timeout 120 python scripts/cta_trace.py flux $OUT/r02c_cta_trace_$V.txt > /dev/null 2>>$OUT/r02c_err.txt
  ( timeout 120 python scripts/fwd_trace.py bf16 2 2>&1 | tail -12 ) > $OUT/r02c_fwd_trace_$V.txt
  python - <<PY
import json
for f in ("flux", "flux_causal"):
    try:
        d=json.load(open("$OUT/r02c_bench_%s_$V.json" % f))
        print("$V", f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"])
    except Exception as e: print("$V", f, "failed", e)
PY
  grep -E "loop|epilogue|first_S|kernel span" $OUT/r02c_cta_trace_$V.txt
  cat $OUT/r02c_fwd_trace_$V.txt
done
export MFA_LIBRARY=$PWD/lib_variants/nostage/libMFAFFI.so
timeout 200 python scripts/ring_emulate.py 131072 32 8 > $OUT/r02c_ring_emulate_nostage.txt 2>>$OUT/r02c_err.txt; cat $OUT/r02c_ring_emulate_nostage.txt
tail -5 $OUT/r02c_err.txt
