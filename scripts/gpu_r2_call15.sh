#!/bin/bash
# round 2: quantised backward on the tensor pipe
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_tcq.py tests/test_gpu_quant.py -m gpu -q --tb=short -k "backward" 2>&1 | cut -c1-300 | tail -40 ) > $OUT/r02p_tcq_bwd_tests.log; cat $OUT/r02p_tcq_bwd_tests.log
