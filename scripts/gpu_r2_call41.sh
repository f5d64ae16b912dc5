#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 400 python -m pytest tests/test_gpu_ring.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -25 ) > $OUT/r02ap_ring_tests.log; cat $OUT/r02ap_ring_tests.log
