#!/bin/bash
# round 2: int8 / int4 scores accumulated onto a float bias in TMEM (tcgen05.cp + tcgen05.st), no per-score widening
OUT=gpurun_out; mkdir -p $OUT
( timeout 240 python -m pytest tests/test_gpu_tcq.py tests/test_gpu_quant.py -m gpu -q --tb=short -x 2>&1 | cut -c1-300 | tail -30 ) > $OUT/r02w_tcq_tests.log; cat $OUT/r02w_tcq_tests.log
timeout 200 python scripts/bench_quant.py 10 > $OUT/r02w_bench_quant.json 2>$OUT/r02w_err.txt
MFA_FWD_POLY=0 timeout 200 python scripts/bench_quant.py 10 > $OUT/r02w_bench_quant_poly0.json 2>>$OUT/r02w_err.txt
MFA_TCQ_PV=bf16 timeout 200 python scripts/bench_quant.py 10 > $OUT/r02w_bench_quant_bf16pv.json 2>>$OUT/r02w_err.txt
python - <<PY
import json
for f in ("", "_poly0", "_bf16pv"):
    try:
        d=json.load(open("$OUT/r02w_bench_quant%s.json" % f))
        print(f or "default", {k: (round(v["ms"],4), round(v.get("cosine_vs_bf16",1),5), v["kernel"]) for k,v in d.items() if isinstance(v, dict)})
    except Exception as e: print(f, "failed", e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02w_launches_quant.csv python scripts/bench_quant.py 2 > /dev/null 2>&1
grep -v "^==" $OUT/r02w_launches_quant.csv | awk -F'","' 'NR>1{print $5, $NF}' | grep -v "at::\|cublas\|dot_k\|reduce_1" | tail -12 | cut -c1-160
tail -3 $OUT/r02w_err.txt
