#!/bin/bash
# round 2: head dims that are not 64 / 128 / 256 on the tensor pipe (zero-padded by TMA), forward and backward
OUT=gpurun_out; mkdir -p $OUT
( timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -40 ) > $OUT/r02z_padded_tests.log; cat $OUT/r02z_padded_tests.log
