#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_fp32_tc.py -m gpu -q --tb=short -k "same_sign" 2>&1 | cut -c1-300 | tail -15 ) > $OUT/r02ak_fp32_same_sign.log; cat $OUT/r02ak_fp32_same_sign.log
( MFA_FP32_SLICE_KEYS=0 timeout 300 python -m pytest tests/test_gpu_fp32_tc.py -m gpu -q --tb=line -k "same_sign" 2>&1 | grep "rel err\|passed\|failed" | cut -c1-200 | tail -5 ) > $OUT/r02ak_fp32_same_sign_noslice.log; cat $OUT/r02ak_fp32_same_sign_noslice.log
( MFA_FP32_SLICE_KEYS=2048 timeout 300 python -m pytest tests/test_gpu_fp32_tc.py -m gpu -q --tb=line -k "same_sign" 2>&1 | grep "rel err\|passed\|failed" | cut -c1-200 | tail -5 ) > $OUT/r02ak_fp32_same_sign_2048.log; cat $OUT/r02ak_fp32_same_sign_2048.log
