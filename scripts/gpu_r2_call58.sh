#!/bin/bash
# block quantiser CTA size A/B (128 vs 256 threads): bit-exact tests with the default, helper launch lists for both, quantised bench
TAG=${1:-r02bo}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_quant.py tests/test_gpu_tcq.py -m gpu -q -x 2>&1 | tail -3 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
for M in 128 256; do
  MFA_QUANT_SPAN_THREADS=$M timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
      --log-file $OUT/${TAG}_helpers_$M.csv python scripts/bench_helpers.py 3 > /dev/null 2>$OUT/${TAG}_err.txt
  echo "threads $M quant_span ns:"; grep "quant_span" $OUT/${TAG}_helpers_$M.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
done
timeout 300 python scripts/bench_quant.py 10 > $OUT/${TAG}_bench_quant.json 2>> $OUT/${TAG}_err.txt
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_quant.json").read().strip().splitlines()[-1])
print({k: (round(v["ms"], 4), round(v.get("speedup_vs_bf16_incl_quantise", 1), 3)) for k, v in d.items() if isinstance(v, dict)})
PY
tail -3 $OUT/${TAG}_err.txt
