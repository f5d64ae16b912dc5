#!/bin/bash
# round 2: fp32 split path after the pre-pass rework (one launch per pass, no index divisions for packed sources)
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_fp32_tc.py tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -15 ) > $OUT/r02v_fp32_tests.log; cat $OUT/r02v_fp32_tests.log
timeout 300 python scripts/bench_fp32.py 10 > $OUT/r02v_bench_fp32.json 2>$OUT/r02v_err.txt; cat $OUT/r02v_bench_fp32.json
timeout 300 python scripts/bench_fp32.py 10 causal > $OUT/r02v_bench_fp32_causal.json 2>>$OUT/r02v_err.txt; cat $OUT/r02v_bench_fp32_causal.json
MFA_FP32_SLICE_KEYS=0 timeout 300 python scripts/bench_fp32.py 10 > $OUT/r02v_bench_fp32_noslice.json 2>>$OUT/r02v_err.txt; cat $OUT/r02v_bench_fp32_noslice.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/r02v_launches_fp32.csv python scripts/bench_fp32.py 2 > /dev/null 2>&1
grep -v "^==" $OUT/r02v_launches_fp32.csv | awk -F'","' 'NR>1{print $5, $NF}' | grep -v "at::\|simt" | head -8 | cut -c1-150
tail -3 $OUT/r02v_err.txt
