#!/bin/bash
# round 2: shifted-bias widening (int4 precision fix) + where the int8 / e4m3 kernel's time goes (timeline, turn-taking off)
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_tcq.py tests/test_gpu_quant.py -m gpu -q --tb=line 2>&1 | grep -v "^  " | cut -c1-300 | tail -30 ) > $OUT/r02o_tcq_tests.log; cat $OUT/r02o_tcq_tests.log
timeout 300 python scripts/bench_quant.py 10 > $OUT/r02o_bench_quant.json 2>$OUT/r02o_err.txt
MFA_FWD_PINGPONG=0 timeout 300 python scripts/bench_quant.py 10 > $OUT/r02o_bench_quant_nopingpong.json 2>>$OUT/r02o_err.txt
MFA_FWD_PINGPONG=0 MFA_FWD_POLY=0 timeout 300 python scripts/bench_quant.py 10 > $OUT/r02o_bench_quant_nopingpong_poly0.json 2>>$OUT/r02o_err.txt
MFA_FWD_POLY=0 timeout 300 python scripts/bench_quant.py 10 > $OUT/r02o_bench_quant_poly0.json 2>>$OUT/r02o_err.txt
python - <<PY
import json
for f in ("", "_nopingpong", "_nopingpong_poly0", "_poly0"):
    try:
        d=json.load(open("$OUT/r02o_bench_quant%s.json" % f))
        print(f or "default", {k: (round(v["ms"],4), round(v.get("cosine_vs_bf16",1),5), v["kernel"]) for k,v in d.items() if isinstance(v, dict)})
    except Exception as e: print(f, "failed", e)
PY
for P in 3 0; do
timeout 120 python scripts/fwd_trace.py int8 $P 2 > $OUT/r02o_trace_int8_poly$P.txt 2>>$OUT/r02o_err.txt
timeout 60 python scripts/fwd_trace_events.py $OUT/fwd_trace_int8_${P}_q2.txt 10 2 >> $OUT/r02o_trace_int8_poly$P.txt 2>>$OUT/r02o_err.txt
done
timeout 120 python scripts/fwd_trace.py bf16 3 > $OUT/r02o_trace_bf16.txt 2>>$OUT/r02o_err.txt
timeout 60 python scripts/fwd_trace_events.py $OUT/fwd_trace_bf16_3_q2.txt 10 2 >> $OUT/r02o_trace_bf16.txt 2>>$OUT/r02o_err.txt
head -8 $OUT/r02o_trace_int8_poly3.txt; head -8 $OUT/r02o_trace_bf16.txt
tail -5 $OUT/r02o_err.txt
