#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 500 python -m pytest tests/test_gpu_ring.py -m gpu -q --tb=short -k "backward" 2>&1 | cut -c1-300 | tail -25 ) > $OUT/r02aq_ring_bwd_tests.log; cat $OUT/r02aq_ring_bwd_tests.log
