#!/bin/bash
# staged (TMA) external masks in the forward: mask tests of the tensor-core forward (watchdog build), mask bench
TAG=${1:-r02ba}
OUT=gpurun_out
mkdir -p $OUT
( timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "mask" 2>&1 | tail -15 ) > $OUT/${TAG}_mask_tests.log
cat $OUT/${TAG}_mask_tests.log
( timeout 400 python -m pytest tests/test_gpu_tc_bwd.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 ) > $OUT/${TAG}_bwd_tests.log; cat $OUT/${TAG}_bwd_tests.log
timeout 200 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2> $OUT/${TAG}_err.txt
MFA_DISABLE_MASK_TMA=1 timeout 200 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask_inplace.json 2>> $OUT/${TAG}_err.txt
python - <<PY
import json
for f in ("bench_mask", "bench_mask_inplace"):
    try:
        d = json.loads(open("$OUT/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: (round(v["ms"], 4), v["kernel"], round(v.get("bwd_ms", 0), 4)) for k, v in d.items() if isinstance(v, dict)})
    except Exception as e: print(f, "failed", e)
PY
tail -5 $OUT/${TAG}_err.txt
