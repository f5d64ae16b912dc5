#!/bin/bash
# Session re-entry visit: parity tests, headline bench, int8 vs bf16 with ping-pong on/off, timelines, ncu full of fwd kernels.
TAG=${1:-v1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_gpu_tests.log
timeout 300 python bench.py > $OUT/${TAG}_bench_flux_fwd.json 2> $OUT/${TAG}_bench.err
for PP in 1 0; do for P in 0 3; do echo "PINGPONG=$PP POLY=$P"; MFA_FWD_PINGPONG=$PP MFA_FWD_POLY=$P timeout 300 python scripts/bench_quant.py 10; done; done > $OUT/${TAG}_quant_sweep.txt 2>&1
( for PP in 1 0; do for m in bf16 int8; do echo "PINGPONG=$PP"; MFA_FWD_PINGPONG=$PP timeout 120 python scripts/fwd_trace.py $m 2 2>&1 | tail -8; done; done ) > $OUT/${TAG}_trace.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_quant.csv \
    python scripts/bench_quant.py 2 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 4 -c 1 \
    -o $OUT/${TAG}_full_fwd_bf16 -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 14 -c 1 \
    -o $OUT/${TAG}_full_fwd_int8 -f python scripts/bench_quant.py 2 > /dev/null 2>&1
ls -la $OUT
cat $OUT/${TAG}_bench.err | tail -5
cut -c1-600 $OUT/${TAG}_bench_flux_fwd.json
cat $OUT/${TAG}_quant_sweep.txt | cut -c1-900
cat $OUT/${TAG}_trace.txt
