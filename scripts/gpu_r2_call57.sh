#!/bin/bash
# closing run of the round on the final build: whole GPU suite, smoke, default bench line with every extra, reference arm
TAG=${1:-r02bn}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_gpu_tests.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > $OUT/${TAG}_smoke.log; cat $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","ms_per_step","gpu_launches") if k in d}, d.get("e2e"), d.get("clocks"), {k: d["roofline"].get(k) for k in ("frac","traffic")}, d.get("cpu_baseline"))
for k,v in (d.get("extras") or {}).items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","scaling","error")}, (v.get("config") or {}).get("kernel"))
try:
    r=json.loads(open("$OUT/${TAG}_bench_reference_arm.json").read().strip().splitlines()[-1]); print("reference_arm", round(r["value"],3), r["unit"], r.get("ms_per_step"))
except Exception as e: print("reference arm failed", e)
PY
tail -3 $OUT/${TAG}_bench.err
