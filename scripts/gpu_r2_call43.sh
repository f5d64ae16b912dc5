#!/bin/bash
# 8-GPU visit: native ring backward -- correctness at world 4, timing at N = 1 and 8 (128k causal)
OUT=gpurun_out; mkdir -p $OUT
( timeout 400 python -m pytest tests/test_gpu_ring.py -m gpu -q --tb=short -k "native_backward" 2>&1 | cut -c1-300 | tail -12 ) > $OUT/r02ar_ring_bwd_tests.log; cat $OUT/r02ar_ring_bwd_tests.log
timeout 400 python scripts/ring_bwd_bench.py > $OUT/r02ar_ring_fwdbwd_n1.json 2>$OUT/r02ar_err.txt; cat $OUT/r02ar_ring_fwdbwd_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 scripts/ring_bwd_bench.py > $OUT/r02ar_ring_fwdbwd_n8.json 2>>$OUT/r02ar_err.txt; cat $OUT/r02ar_ring_fwdbwd_n8.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29578 scripts/ring_bwd_bench.py > $OUT/r02ar_ring_fwdbwd_n4.json 2>>$OUT/r02ar_err.txt; cat $OUT/r02ar_ring_fwdbwd_n4.json
tail -3 $OUT/r02ar_err.txt | cut -c1-300
