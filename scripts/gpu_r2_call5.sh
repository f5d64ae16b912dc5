#!/bin/bash
# round 2, GPU visit 5: forward v2b (strict exp2 turn-taking, warp-local coalesced epilogue, epilogue warps lowest priority)
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/r02f_gpu_tests.log; cat $OUT/r02f_gpu_tests.log
timeout 200 python scripts/ring_emulate.py 131072 32 8 > $OUT/r02f_ring_emulate.txt 2>$OUT/r02f_ring_err.txt; cat $OUT/r02f_ring_emulate.txt; grep "mfa\]" $OUT/r02f_ring_err.txt | head -30
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
for W in flux flux_causal long_dense; do
  timeout 200 $B --workload $W > $OUT/r02f_bench_$W.json 2>>$OUT/r02f_err.txt
done
MFA_FWD_PERSIST=0 timeout 200 $B > $OUT/r02f_bench_flux_nopersist.json 2>>$OUT/r02f_err.txt
timeout 120 python scripts/cta_trace.py flux $OUT/r02f_cta_trace.txt > /dev/null 2>>$OUT/r02f_err.txt
( timeout 120 python scripts/fwd_trace.py bf16 2 2>&1 | tail -8 ) > $OUT/r02f_fwd_trace.txt
python - <<PY
import json
for f in ("flux", "flux_causal", "long_dense", "flux_nopersist"):
    try:
        d=json.load(open("$OUT/r02f_bench_%s.json" % f))
        print(f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
grep -E "loop|epilogue|first_S|pv_tail|kernel span" $OUT/r02f_cta_trace.txt; cat $OUT/r02f_fwd_trace.txt; grep "mfa\]" $OUT/r02f_err.txt | head; tail -3 $OUT/r02f_err.txt
