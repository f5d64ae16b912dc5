#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tcq.py -x -q 2>&1 | tail -15 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
( bash scripts/poly_sweep.sh ) > $OUT/${TAG}_poly.txt 2>&1
cat $OUT/${TAG}_poly.txt
for P in 0 2; do echo "quant POLY=$P"; MFA_FWD_POLY=$P timeout 300 python scripts/bench_quant.py 10; done > $OUT/${TAG}_quant.txt 2>&1
cat $OUT/${TAG}_quant.txt
( for m in bf16 int8; do for P in 0 2; do timeout 120 python scripts/fwd_trace.py $m $P 2>&1 | tail -8; done; done ) > $OUT/${TAG}_trace.txt 2>&1
cat $OUT/${TAG}_trace.txt
