#!/bin/bash
# round 2: int8 Q K^T without I2FP + e4m3 P V
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_tcq.py tests/test_gpu_quant.py tests/test_gpu_torch_adapter.py -m gpu -q 2>&1 | tail -25 ) > $OUT/r02m_tcq_tests.log; cat $OUT/r02m_tcq_tests.log
timeout 300 python scripts/bench_quant.py 10 > $OUT/r02m_bench_quant_fp8.json 2>$OUT/r02m_err.txt
MFA_TCQ_PV=bf16 timeout 300 python scripts/bench_quant.py 10 > $OUT/r02m_bench_quant_bf16pv.json 2>>$OUT/r02m_err.txt
for P in 0 2 4; do MFA_FWD_POLY=$P timeout 300 python scripts/bench_quant.py 10 > $OUT/r02m_bench_quant_fp8_poly$P.json 2>>$OUT/r02m_err.txt; done
python - <<PY
import json
for f in ("fp8", "bf16pv", "fp8_poly0", "fp8_poly2", "fp8_poly4"):
    try:
        d=json.load(open("$OUT/r02m_bench_quant_%s.json" % f))
        print(f, {k: (round(v["ms"],4), round(v.get("cosine_vs_bf16",1),5), v["kernel"]) for k,v in d.items() if isinstance(v, dict)})
    except Exception as e: print(f, "failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02m_launches_quant.csv python scripts/bench_quant.py 2 > /dev/null 2>&1
grep -v "^==" $OUT/r02m_launches_quant.csv | awk -F'","' 'NR>1{print $5, $NF}' | tail -24
tail -3 $OUT/r02m_err.txt
