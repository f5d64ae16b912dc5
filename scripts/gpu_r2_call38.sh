#!/bin/bash
# round 2: fp32 split path with the in-kernel O flush (one launch) vs one launch per slice
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_fp32_tc.py tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -12 ) > $OUT/r02am_fp32_tests.log; cat $OUT/r02am_fp32_tests.log
timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02am_bench_fp32_flush.json 2>$OUT/r02am_err.txt; cut -c1-330 $OUT/r02am_bench_fp32_flush.json; echo
MFA_FP32_SLICE_LAUNCHES=1 timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02am_bench_fp32_launches.json 2>>$OUT/r02am_err.txt; cut -c1-330 $OUT/r02am_bench_fp32_launches.json; echo
timeout 200 python scripts/bench_fp32.py 10 causal > $OUT/r02am_bench_fp32_flush_causal.json 2>>$OUT/r02am_err.txt; cut -c1-330 $OUT/r02am_bench_fp32_flush_causal.json; echo
tail -2 $OUT/r02am_err.txt | cut -c1-200
