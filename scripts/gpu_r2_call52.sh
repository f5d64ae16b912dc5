#!/bin/bash
# warp-level Hadamard kernel + small-CTA mask pre-pass: quantiser / mask tests, helper kernels under ncu (time + DRAM bytes per launch),
# mask bench, launch list of the masked forward, ncu --set full of the headline forward of these sources (roofline.traffic)
TAG=${1:-r02bg}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_quant.py tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py tests/test_gpu_torch_adapter.py -m gpu -q -x 2>&1 | tail -6 ) > $OUT/${TAG}_tests.log
cat $OUT/${TAG}_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file $OUT/${TAG}_helpers.csv python scripts/bench_helpers.py 3 > /dev/null 2>$OUT/${TAG}_err.txt
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/${TAG}_helpers.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][-60:]
    agg.setdefault((name, r[-3]), []).append(float(r[-1].replace(",", "")))
for (name, metric), v in agg.items():
    if "at::" in name or "distribution" in name: continue
    print(f"{name:62s} {metric:26s} n={len(v):2d} median={sorted(v)[len(v)//2]:.3f}")
PY
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2>> $OUT/${TAG}_err.txt
MFA_BENCH_MASK_FWD_ONLY=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_mask.csv \
    python scripts/bench_mask.py 2 > /dev/null 2>&1
grep "mask_flags\|mask_compact" $OUT/${TAG}_launches_mask.csv | awk -F'","' '{print $5, $NF}' | sed -E 's/\(.*\)//' | tr '\n' ' '; echo
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
tail -3 $OUT/${TAG}_err.txt
