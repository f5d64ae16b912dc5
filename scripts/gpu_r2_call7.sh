#!/bin/bash
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/r02h_gpu_tests.log; cat $OUT/r02h_gpu_tests.log
timeout 300 python scripts/ring_emulate_single.py 131072 32 1,8 8 > $OUT/r02h_ring_single.txt 2>$OUT/r02h_err.txt; cat $OUT/r02h_ring_single.txt
timeout 300 python scripts/ring_emulate_single.py 131072 32 8 0 >> $OUT/r02h_ring_single.txt 2>>$OUT/r02h_err.txt; tail -2 $OUT/r02h_ring_single.txt
timeout 300 python bench.py --workload ring128k --steps 3 --warmup 2 > $OUT/r02h_bench_ring128k_n1.json 2>>$OUT/r02h_err.txt; cut -c1-300 $OUT/r02h_bench_ring128k_n1.json
grep "mfa\]" $OUT/r02h_err.txt | head -20; tail -5 $OUT/r02h_err.txt
