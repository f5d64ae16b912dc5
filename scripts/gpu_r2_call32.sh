#!/bin/bash
# round 2: 2-GPU visit -- ring tests over both transports with the final kernels + ring128k / C4 strong scaling at N=2
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_ring.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -15 ) > $OUT/r02ag_ring_tests_2gpu.log; cat $OUT/r02ag_ring_tests_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > $OUT/r02ag_bench_n2.json 2>$OUT/r02ag_err.txt
python - <<PY
import json
d=json.loads(open("$OUT/r02ag_bench_n2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value","n_gpus","ms_per_step","scaling")}, d.get("e2e"))
for k,v in (d.get("extras") or {}).items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","scaling","n_gpus","error")})
PY
tail -3 $OUT/r02ag_err.txt | cut -c1-300
