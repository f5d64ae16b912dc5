#!/bin/bash
# round 2, 8-GPU visit with the final kernels: bench.py exactly as the driver launches it (N = 1 and N = 8, every extra)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r02an_topo.txt 2>&1
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/r02an_bench_n1.json 2>$OUT/r02an_err.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/r02an_bench_n8.json 2>>$OUT/r02an_err.txt
MFA_RING_TRANSPORT=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --workload ring128k --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r02an_ring128k_n8_nccl.json 2>>$OUT/r02an_err.txt
python - <<PY
import json
base={}
for f in ("n1","n8"):
    try:
        d=json.loads(open("$OUT/r02an_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value","n_gpus","ms_per_step","scaling")}, "e2e", round(d["e2e"]["value"],1), d.get("clocks"))
        for k,v in (d.get("extras") or {}).items():
            val=v.get("value"); 
            if f=="n1": base[k]=val
            eff = (val/base[k]/(8 if v.get("scaling")=="weak" or True else 1)) if (f=="n8" and base.get(k) and val) else None
            print("  ", k, {x: v.get(x) for x in ("value","ms_per_step","scaling","error")}, ("eff vs 8x N=1: %.3f" % eff) if eff else "")
    except Exception as e: print(f, "failed", e)
try:
    d=json.loads(open("$OUT/r02an_ring128k_n8_nccl.json").read().strip().splitlines()[-1]); print("ring128k n8 nccl", round(d["value"],1), round(d["ms_per_step"],3))
except Exception as e: print("nccl failed", e)
PY
tail -3 $OUT/r02an_err.txt | cut -c1-300
