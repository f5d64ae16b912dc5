#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -30 ) > $OUT/r02aa_gpu_tests.log; cat $OUT/r02aa_gpu_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --extras none > $OUT/r02aa_bench_flux.json 2>$OUT/r02aa_err.txt; cut -c1-200 $OUT/r02aa_bench_flux.json; echo
timeout 200 python bench.py --workload ring128k --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $OUT/r02aa_bench_ring1.json 2>>$OUT/r02aa_err.txt; cut -c1-200 $OUT/r02aa_bench_ring1.json; echo
tail -3 $OUT/r02aa_err.txt
