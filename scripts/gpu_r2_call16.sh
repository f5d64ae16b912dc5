#!/bin/bash
# round 2: fp32 operands on the tensor pipe (fp16 hi/lo split)
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_fp32_tc.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -40 ) > $OUT/r02q_fp32_tests.log; cat $OUT/r02q_fp32_tests.log
timeout 300 python scripts/bench_fp32.py 10 > $OUT/r02q_bench_fp32.json 2>$OUT/r02q_err.txt; cat $OUT/r02q_bench_fp32.json; tail -3 $OUT/r02q_err.txt
