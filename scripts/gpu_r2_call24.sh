#!/bin/bash
# round 2: head_dim 256 on the tensor pipe (wide mode)
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x -k "d256" 2>&1 | cut -c1-300 | tail -30 ) > $OUT/r02y_d256_tests.log; cat $OUT/r02y_d256_tests.log
timeout 200 python bench.py --workload d256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --extras none > $OUT/r02y_bench_d256.json 2>$OUT/r02y_err.txt
cut -c1-600 $OUT/r02y_bench_d256.json; tail -3 $OUT/r02y_err.txt
