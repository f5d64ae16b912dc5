#!/bin/bash
# round 2, GPU visit 1: baseline sanity + diagnostics (per-CTA timeline, fp8 layout probe, persistent / long-run variants)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/r02a_gpu.txt 2>&1
./tools/f8_probe > $OUT/r02a_f8_probe.txt 2>&1; cat $OUT/r02a_f8_probe.txt
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $OUT/r02a_gpu_tests.log; cat $OUT/r02a_gpu_tests.log
timeout 120 python scripts/cta_trace.py flux $OUT/r02a_cta_trace_flux.txt 2>&1 | tail -20
timeout 120 python scripts/cta_trace.py flux_causal $OUT/r02a_cta_trace_flux_causal.txt 2>&1 | tail -20
B="python bench.py --no-cpu-baseline --no-e2e"
timeout 200 $B > $OUT/r02a_bench_flux.json 2>$OUT/r02a_err.txt
MFA_FWD_PERSIST=1 timeout 200 $B > $OUT/r02a_bench_flux_persist.json 2>>$OUT/r02a_err.txt
timeout 200 $B --steps 400 --warmup 50 > $OUT/r02a_bench_flux_400.json 2>>$OUT/r02a_err.txt
MFA_DEBUG_SKIP_STORE=1 timeout 200 $B > $OUT/r02a_bench_flux_nostore.json 2>>$OUT/r02a_err.txt
timeout 200 $B --workload long_dense --steps 5 --warmup 2 > $OUT/r02a_bench_long_dense.json 2>>$OUT/r02a_err.txt
timeout 200 python scripts/ring_emulate.py 131072 32 8 > $OUT/r02a_ring_emulate.txt 2>>$OUT/r02a_err.txt
for f in flux flux_persist flux_400 flux_nostore long_dense; do python - <<PY
import json
d=json.load(open("$OUT/r02a_bench_$f.json"))
print("$f", round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"])
PY
done
cat $OUT/r02a_ring_emulate.txt; tail -3 $OUT/r02a_err.txt
