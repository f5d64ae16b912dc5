#!/bin/bash
# round 2: full GPU suite + smoke + default bench after the fp32 split path / e4m3 P V / quantised TC backward
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -40 ) > $OUT/r02r_gpu_tests.log; cat $OUT/r02r_gpu_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $OUT/r02r_smoke.log; cat $OUT/r02r_smoke.log
( time timeout 900 python bench.py > $OUT/r02r_bench_default.json 2>$OUT/r02r_bench_err.txt ) 2>&1 | tail -3
python - <<PY
import json
d=json.loads(open("$OUT/r02r_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","ms_per_step","e2e","roofline","gpu_launches","clocks") if k in d})
for k,v in (d.get("extras") or {}).items(): print(k, json.dumps(v)[:400])
PY
tail -3 $OUT/r02r_bench_err.txt
