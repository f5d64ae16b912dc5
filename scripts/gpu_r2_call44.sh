#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras ring128k,ring128k_fwdbwd > $OUT/r02as_bench_ring_n1.json 2>$OUT/r02as_err.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras ring128k,ring128k_fwdbwd > $OUT/r02as_bench_ring_n2.json 2>>$OUT/r02as_err.txt
python - <<PY
import json
for f in ("n1","n2"):
    try:
        d=json.loads(open("$OUT/r02as_bench_ring_%s.json" % f).read().strip().splitlines()[-1])
        for k,v in (d.get("extras") or {}).items(): print(f, k, {x: v.get(x) for x in ("metric","value","ms_per_step","scaling","n_gpus","error","gpu_launches")})
    except Exception as e: print(f, "failed", e)
PY
tail -3 $OUT/r02as_err.txt | cut -c1-300
