#!/bin/bash
# round 2, GPU visit 2: forward v2 (persistent + epilogue warpgroup) -- parity tests, then timings
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/r02b_gpu_tests.log; cat $OUT/r02b_gpu_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
timeout 200 $B > $OUT/r02b_bench_flux.json 2>$OUT/r02b_err.txt
MFA_FWD_PERSIST=0 timeout 200 $B > $OUT/r02b_bench_flux_nopersist.json 2>>$OUT/r02b_err.txt
timeout 200 $B --workload flux_causal > $OUT/r02b_bench_flux_causal.json 2>>$OUT/r02b_err.txt
timeout 200 $B --workload long_dense --steps 5 --warmup 2 > $OUT/r02b_bench_long_dense.json 2>>$OUT/r02b_err.txt
timeout 200 $B --mode fwdbwd > $OUT/r02b_bench_flux_fwdbwd.json 2>>$OUT/r02b_err.txt
timeout 200 python scripts/ring_emulate.py 131072 32 8 > $OUT/r02b_ring_emulate.txt 2>>$OUT/r02b_err.txt
timeout 120 python scripts/cta_trace.py flux $OUT/r02b_cta_trace_flux.txt > /dev/null 2>>$OUT/r02b_err.txt
for f in flux flux_nopersist flux_causal long_dense flux_fwdbwd; do python - <<PY
import json
try:
    d=json.load(open("$OUT/r02b_bench_$f.json"))
    print("$f", round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"])
except Exception as e: print("$f failed", e)
PY
done
cat $OUT/r02b_ring_emulate.txt; cat $OUT/r02b_cta_trace_flux.txt; tail -5 $OUT/r02b_err.txt
