#!/bin/bash
# Round-2 evidence run on one B200: parity tests, smoke, the default bench line with every extra, both arms, per-mode benches,
# ncu launch lists.   gpurun --timeout 1700 -- 'bash scripts/gpu_final_r2.sh r02f'
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_gpu_tests.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > $OUT/${TAG}_smoke.log; cat $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload flux_causal --no-cpu-baseline --no-e2e --extras none > $OUT/${TAG}_bench_flux_causal.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload d256 --steps 10 --no-cpu-baseline --no-e2e --extras none > $OUT/${TAG}_bench_d256.json 2>> $OUT/${TAG}_bench.err
timeout 300 python scripts/bench_quant.py 10 > $OUT/${TAG}_bench_quant.json 2>> $OUT/${TAG}_bench.err
timeout 300 python scripts/bench_fp32.py 10 > $OUT/${TAG}_bench_fp32.json 2>> $OUT/${TAG}_bench.err
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2>> $OUT/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras fwdbwd_flux,int8_block,int4_block,fp32_flux,d256_fwd > /dev/null 2>&1
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","ms_per_step","gpu_launches") if k in d}, d.get("e2e"), d.get("clocks"))
for k,v in (d.get("extras") or {}).items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","scaling","error")}, (v.get("config") or {}).get("kernel"))
for f in ("reference_arm","flux_causal","d256"):
    try:
        r=json.loads(open("$OUT/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1]); print(f, round(r["value"],3), r["unit"], r.get("ms_per_step"))
    except Exception as e: print(f, "failed", e)
for f in ("quant","fp32","mask"):
    try: print(f, open("$OUT/${TAG}_bench_%s.json" % f).read()[:900])
    except Exception as e: print(f, "failed", e)
PY
tail -3 $OUT/${TAG}_bench.err
