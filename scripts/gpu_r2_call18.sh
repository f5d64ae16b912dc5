#!/bin/bash
# round 2: int4 Q / K unpacked in shared memory by a converter warp (no global int4 -> int8 pass)
OUT=gpurun_out; mkdir -p $OUT
( timeout 240 python -m pytest tests/test_gpu_tcq.py tests/test_gpu_quant.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -40 ) > $OUT/r02s_tcq_tests.log; cat $OUT/r02s_tcq_tests.log
timeout 300 python scripts/bench_quant.py 10 > $OUT/r02s_bench_quant.json 2>$OUT/r02s_err.txt
python - <<PY
import json
d=json.load(open("$OUT/r02s_bench_quant.json"))
print({k: (round(v["ms"],4), round(v.get("cosine_vs_bf16",1),5), v["kernel"]) for k,v in d.items() if isinstance(v, dict)})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02s_launches_quant.csv python scripts/bench_quant.py 2 > /dev/null 2>&1
grep -v "^==" $OUT/r02s_launches_quant.csv | awk -F'","' 'NR>1{print $5, $NF}' | tail -14 | cut -c1-200
tail -3 $OUT/r02s_err.txt
