#!/bin/bash
# round 2, 2-GPU visit: native ring over NCCL and over copy engines; single-GPU sanity after restoring the round-1 forward structure
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r02k_topo.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/r02k_gpu_tests.log; cat $OUT/r02k_gpu_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --extras none"
for W in flux flux_causal; do timeout 200 $B --workload $W > $OUT/r02k_bench_$W.json 2>>$OUT/r02k_err.txt; done
timeout 300 $B --workload ring128k --steps 3 --warmup 2 > $OUT/r02k_ring128k_n1.json 2>>$OUT/r02k_err.txt
for T in nccl p2p; do
  MFA_RING_TRANSPORT=$T MFA_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --workload ring128k --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $OUT/r02k_ring128k_n2_$T.json 2>$OUT/r02k_ring_n2_$T.err
  tail -1 $OUT/r02k_ring128k_n2_$T.json | cut -c1-200; grep -i "mfa\|error" $OUT/r02k_ring_n2_$T.err | head -5
done
python - <<PY
import json
for f in ("bench_flux", "bench_flux_causal", "ring128k_n1", "ring128k_n2_nccl", "ring128k_n2_p2p"):
    try:
        d=json.loads(open("$OUT/r02k_%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "TFLOP/s", round(d["ms_per_step"],4), "ms", d["clocks"]["sm_mhz"], d["config"].get("parallelism", "")[:90])
    except Exception as e: print(f, "failed", e)
PY
tail -3 $OUT/r02k_err.txt
