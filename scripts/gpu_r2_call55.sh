#!/bin/bash
# mask pre-pass chosen by tile count: mask / backward tests, mask bench + launch list, headline capture of the final sources, smoke
TAG=${1:-r02bk}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py tests/test_gpu_tcq.py tests/test_gpu_fp32_tc.py tests/test_gpu_torch_adapter.py -m gpu -q -x 2>&1 | tail -4 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > $OUT/${TAG}_smoke.log; cat $OUT/${TAG}_smoke.log
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2> $OUT/${TAG}_err.txt
MFA_BENCH_MASK_FWD_ONLY=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_mask.csv \
    python scripts/bench_mask.py 2 > /dev/null 2>&1
echo "mask_flags ns:"; grep "mask_flags" $OUT/${TAG}_launches_mask.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_mask.json").read().strip().splitlines()[-1])
    print({k: (round(v["ms"], 4), round(v.get("bwd_ms", 0), 4)) for k, v in d.items() if isinstance(v, dict)})
except Exception as e: print("mask failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 4 -c 1 -o $OUT/${TAG}_full_fwd_bf16 -f \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --extras none > /dev/null 2>>$OUT/${TAG}_err.txt
timeout 200 python scripts/ncu_summary.py $OUT/${TAG}_full_fwd_bf16.ncu-rep 12 > $OUT/${TAG}_ncu_fwd_bf16.txt 2>&1
timeout 100 ncu -i $OUT/${TAG}_full_fwd_bf16.ncu-rep --page details --csv 2>/dev/null | grep -i "pipe\|Executed Ipc\|Issue Slots\|Duration\|DRAM Throughput\|Registers\|Theoretical Occ\|Memory Throughput" | cut -c1-220 >> $OUT/${TAG}_ncu_fwd_bf16.txt
head -4 $OUT/${TAG}_ncu_fwd_bf16.txt | cut -c1-160
rm -f $OUT/*.ncu-rep
timeout 300 python bench.py --steps 20 --no-cpu-baseline --extras fwdbwd_flux,mask_bf16_dense > $OUT/${TAG}_bench_short.json 2>> $OUT/${TAG}_err.txt
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_short.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"].get("traffic"), d["roofline"].get("traffic_note"))
for k,v in (d.get("extras") or {}).items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","error")})
PY
tail -3 $OUT/${TAG}_err.txt
