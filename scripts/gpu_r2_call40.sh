#!/bin/bash
# same-box A/B: split TMEM read-out of S (HEAD working tree) vs the committed kernel
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fp32_tc.py tests/test_gpu_tcq.py -m gpu -q --tb=short 2>&1 | tail -3 ) ; 
for i in 1 2 3; do
  for v in prev head; do
    if [ $v = head ]; then L=""; else L="MFA_LIBRARY=$PWD/lib_variants/$v/libMFAFFI.so"; fi
    env $L timeout 200 python scripts/bench_quant.py 10 > $OUT/r02ao_${v}_quant_${i}.json 2>>$OUT/r02ao_err.txt
    for wl in flux flux_causal d256 flux_fp32; do
      env $L timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --extras none > $OUT/r02ao_${v}_${wl}_${i}.json 2>>$OUT/r02ao_err.txt
    done
  done
done
python - <<PY
import json
for wl in ("flux","flux_causal","d256","flux_fp32"):
    for v in ("prev","head"):
        vals=[]
        for i in (1,2,3):
            try:
                d=json.loads(open("$OUT/r02ao_%s_%s_%d.json" % (v,wl,i)).read().strip().splitlines()[-1]); vals.append(round(d["value"],1))
            except Exception as e: vals.append(None)
        print(wl, v, vals)
PY
python - <<PY
import json
for v in ("prev","head"):
    for i in (1,2,3):
        try:
            d=json.load(open("$OUT/r02ao_%s_quant_%d.json" % (v,i))); print("quant", v, i, {k: round(x["ms"],4) for k,x in d.items() if isinstance(x, dict)})
        except Exception as e: print("quant", v, i, "failed", e)
PY
tail -2 $OUT/r02ao_err.txt
