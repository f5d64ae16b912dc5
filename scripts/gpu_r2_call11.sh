#!/bin/bash
# round 2, 8-GPU visit: ring128k strong scaling, both transports (N=1 baseline on the same box first)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r02l_topo.txt 2>&1
B="bench.py --workload ring128k --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 python $B > $OUT/r02l_ring128k_n1.json 2>$OUT/r02l_err.txt
for T in p2p nccl; do
  for N in 8 4; do
  MFA_RING_TRANSPORT=$T timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
     $B --gpus $N > $OUT/r02l_ring128k_n${N}_$T.json 2>$OUT/r02l_ring_n${N}_$T.err
  done
done
python - <<PY
import json
base=None
for f in ("ring128k_n1", "ring128k_n8_p2p", "ring128k_n4_p2p", "ring128k_n8_nccl", "ring128k_n4_nccl"):
    try:
        d=json.loads(open("$OUT/r02l_%s.json" % f).read().strip().splitlines()[-1])
        if base is None: base=d["value"]
        print(f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],3), "ms", "eff %.3f" % (d["value"]/base/d["n_gpus"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 $OUT/r02l_err.txt; tail -3 $OUT/r02l_ring_n8_p2p.err
