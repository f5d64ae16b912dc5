#!/bin/bash
# staged masks in the backward kernels (watchdog build): backward + forward mask tests, mask bench, dterm variants under ncu
TAG=${1:-r02be}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_tc_bwd.py tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -12 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2> $OUT/${TAG}_err.txt
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_mask.json").read().strip().splitlines()[-1])
    print({k: (round(v["ms"], 4), round(v.get("bwd_ms", 0), 4), v.get("bwd_kernel")) for k, v in d.items() if isinstance(v, dict)})
except Exception as e: print("mask failed", e)
PY
for R in 1 2 4; do
  MFA_DTERM_ROWS=$R timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches_dterm$R.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras fwdbwd_flux > /dev/null 2>&1
  echo "dterm rows/warp $R:"; grep -i dterm $OUT/${TAG}_launches_dterm$R.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
done
tail -3 $OUT/${TAG}_err.txt
