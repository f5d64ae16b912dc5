#!/bin/bash
# round 2: compute-sanitizer memcheck over the new kernel modes (fp32 split, wide D=256, padded head dims, int4 converter, e4m3 P V, quantised backward)
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fp32_tc.py tests/test_gpu_tcq.py tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py \
    -m gpu -x -q -k "(fp32_tc_shapes or fp32_tc_causal or fp32_tc_external or below_128 or d256 or padded or tcq_int4 or tcq_int8_causal or tcq_backward or additive_mask) and not flux and not saves_time" 2>&1 | grep -v "^=========     at\|^=========         by\|Host Frame" | tail -25 ) > $OUT/r02ae_memcheck.txt
cat $OUT/r02ae_memcheck.txt | cut -c1-200
