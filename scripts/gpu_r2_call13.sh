#!/bin/bash
# round 2: details of the tcq failures (fp8 P V) + full GPU suite state
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_tcq.py -m gpu -q --tb=line 2>&1 | grep -v "^  " | cut -c1-400 | tail -60 ) > $OUT/r02n_tcq_tests.log; cat $OUT/r02n_tcq_tests.log
( timeout 1500 python -m pytest tests -m gpu -q --tb=line --deselect tests/test_gpu_tcq.py 2>&1 | cut -c1-300 | tail -30 ) > $OUT/r02n_gpu_tests.log; cat $OUT/r02n_gpu_tests.log
