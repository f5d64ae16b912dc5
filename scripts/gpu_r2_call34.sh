#!/bin/bash
# round 2: leaner bit-exact quantiser code path + 4-rows-per-warp D-term kernel
OUT=gpurun_out; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_quant.py tests/test_gpu_tcq.py tests/test_gpu_tc_bwd.py tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -12 ) > $OUT/r02ai_tests.log; cat $OUT/r02ai_tests.log
timeout 200 python scripts/bench_quant.py 10 > $OUT/r02ai_bench_quant.json 2>$OUT/r02ai_err.txt
python - <<PY
import json
d=json.load(open("$OUT/r02ai_bench_quant.json"))
print({k: (round(v["ms"],4), round(v.get("cosine_vs_bf16",1),5)) for k,v in d.items() if isinstance(v, dict)})
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $OUT/r02ai_launches_quant.csv python scripts/bench_quant.py 1 > /dev/null 2>&1
grep -v "^==" $OUT/r02ai_launches_quant.csv | awk -F'","' 'NR>1 && ($5 ~ /quant_span|quant_flat|absmax|codes_to|head_v/) {print $5, $(NF-2), $NF}' | cut -c1-40,150-260 | tail -24
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $OUT/r02ai_launches_bwd.csv python bench.py --mode fwdbwd --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras none > /dev/null 2>&1
grep -v "^==" $OUT/r02ai_launches_bwd.csv | awk -F'","' 'NR>1 && ($5 ~ /dterm/) {print $5, $(NF-2), $NF}' | cut -c1-60,100-200 | tail -9
tail -2 $OUT/r02ai_err.txt
