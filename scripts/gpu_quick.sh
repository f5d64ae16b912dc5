#!/bin/bash
# Quick GPU visit for forward-kernel work: TC parity tests, POLY sweep, quantised bench.  Output in gpurun_out/<tag>_*.
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tcq.py -x -q 2>&1 | tail -15 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
( bash scripts/poly_sweep.sh ) > $OUT/${TAG}_poly.txt 2>&1
cat $OUT/${TAG}_poly.txt
for P in 0 2 4; do echo "quant POLY=$P"; MFA_FWD_POLY=$P timeout 300 python scripts/bench_quant.py 10; done > $OUT/${TAG}_quant.txt 2>&1
cat $OUT/${TAG}_quant.txt
timeout 300 python bench.py --workload flux_causal --no-cpu-baseline --no-e2e > $OUT/${TAG}_causal.json 2>&1
timeout 300 python bench.py --workload long_window --steps 10 --no-cpu-baseline --no-e2e > $OUT/${TAG}_c4fwd.json 2>&1
cat $OUT/${TAG}_causal.json $OUT/${TAG}_c4fwd.json | cut -c1-200
