#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_fp32_tc.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -6 ) > $OUT/r02al_fp32_tests.log; cat $OUT/r02al_fp32_tests.log
timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02al_bench_fp32.json 2>$OUT/r02al_err.txt; cut -c1-330 $OUT/r02al_bench_fp32.json; echo
MFA_FP32_SLICE_KEYS=1024 timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02al_bench_fp32_1024.json 2>>$OUT/r02al_err.txt; cut -c1-330 $OUT/r02al_bench_fp32_1024.json; echo
