#!/bin/bash
# staged-mask forward: A/B of library variants (lib_variants/<name>), forward-only mask bench
TAG=${1:-r02bb}; shift
OUT=gpurun_out
mkdir -p $OUT
for V in "$@"; do
  MFA_BENCH_MASK_FWD_ONLY=1 MFA_LIBRARY=lib_variants/$V/libMFAFFI.so timeout 200 python scripts/bench_mask.py 10 > $OUT/${TAG}_mask_$V.json 2> $OUT/${TAG}_err_$V.txt
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_mask_$V.json").read().strip().splitlines()[-1])
    print("$V", {k.split("_[")[0] + k[k.find("["):] if "[" in k else k: round(v["ms"], 4) for k, v in d.items() if isinstance(v, dict)})
except Exception as e: print("$V", "failed", e)
PY
  tail -2 $OUT/${TAG}_err_$V.txt
done
