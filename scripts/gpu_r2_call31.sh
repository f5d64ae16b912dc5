#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02af_fp32_default.json 2>$OUT/r02af_err.txt; cut -c1-260 $OUT/r02af_fp32_default.json; echo
MFA_FWD_PERSIST=1 timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02af_fp32_persist.json 2>>$OUT/r02af_err.txt; cut -c1-260 $OUT/r02af_fp32_persist.json; echo
MFA_FWD_PERSIST=1 MFA_FP32_SLICE_KEYS=0 timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02af_fp32_persist_noslice.json 2>>$OUT/r02af_err.txt; cut -c1-260 $OUT/r02af_fp32_persist_noslice.json; echo
MFA_FP32_SLICE_KEYS=512 timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02af_fp32_512.json 2>>$OUT/r02af_err.txt; cut -c1-260 $OUT/r02af_fp32_512.json; echo
MFA_FWD_PERSIST=1 MFA_FP32_SLICE_KEYS=512 timeout 200 python scripts/bench_fp32.py 10 > $OUT/r02af_fp32_512_persist.json 2>>$OUT/r02af_err.txt; cut -c1-260 $OUT/r02af_fp32_512_persist.json; echo
tail -2 $OUT/r02af_err.txt | cut -c1-200
