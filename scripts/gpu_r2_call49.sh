#!/bin/bash
# release build with TMA-staged forward masks + 4-row dterm: whole GPU suite, mask bench, default bench with the new mask extra
TAG=${1:-r02bd}
OUT=gpurun_out
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_gpu_tests.log
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2> $OUT/${TAG}_err.txt
timeout 400 python bench.py --no-cpu-baseline --extras fwdbwd_flux,mask_bf16_dense > $OUT/${TAG}_bench_default.json 2>> $OUT/${TAG}_err.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_fwdbwd.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras fwdbwd_flux > /dev/null 2>&1
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_mask.json").read().strip().splitlines()[-1])
    print({k: (round(v["ms"], 4), v["kernel"], round(v.get("bwd_ms", 0), 4)) for k, v in d.items() if isinstance(v, dict)})
except Exception as e: print("mask failed", e)
try:
    d=json.loads(open("$OUT/${TAG}_bench_default.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("metric","value","ms_per_step","gpu_launches") if k in d}, d.get("e2e"), d.get("clocks"))
    for k,v in (d.get("extras") or {}).items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","error","mask_hbm")}, (v.get("config") or {}).get("kernel"))
except Exception as e: print("bench failed", e)
PY
grep -i dterm $OUT/${TAG}_launches_fwdbwd.csv | head -3
tail -3 $OUT/${TAG}_err.txt
