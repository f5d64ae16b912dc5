#!/bin/bash
# round 2: e2e host-buffer pipeline with a tapered tail (small last groups) vs uniform groups, and chunk sizes
OUT=gpurun_out; mkdir -p $OUT
B="bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras none"
for v in "tapered16:" "uniform16:MFA_PIPELINE_UNIFORM=1" "tapered24:MFA_PIPELINE_CHUNK_MB=24" "tapered11:MFA_PIPELINE_CHUNK_MB=11" "tapered8:MFA_PIPELINE_CHUNK_MB=8"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 200 python $B > $OUT/r02ah_e2e_$name.json 2>>$OUT/r02ah_err.txt
  python - <<PY
import json
d=json.loads(open("$OUT/r02ah_e2e_$name.json").read().strip().splitlines()[-1])
print("$name", "e2e", round(d["e2e"]["value"],1), "device", round(d["value"],1))
PY
done
( timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2 )
tail -2 $OUT/r02ah_err.txt
