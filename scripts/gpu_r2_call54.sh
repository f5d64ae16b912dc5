#!/bin/bash
# Final evidence run of the round (release build): whole GPU suite, smoke, default bench with every extra, reference arm, causal FLUX, mask
# bench + its launch list, helper launch list, launch list of the default bench, ncu --set full of the headline forward (roofline.traffic)
TAG=${1:-r02bj}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_gpu_tests.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > $OUT/${TAG}_smoke.log; cat $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload flux_causal --no-cpu-baseline --no-e2e --extras none > $OUT/${TAG}_bench_flux_causal.json 2>> $OUT/${TAG}_bench.err
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2>> $OUT/${TAG}_bench.err
MFA_BENCH_MASK_FWD_ONLY=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_mask.csv \
    python scripts/bench_mask.py 2 > /dev/null 2>&1
echo "mask_flags us:"; grep "mask_flags" $OUT/${TAG}_launches_mask.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file $OUT/${TAG}_helpers.csv python scripts/bench_helpers.py 3 > /dev/null 2>>$OUT/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras fwdbwd_flux,mask_bf16_dense,int8_block,fp32_flux,d256_fwd > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 4 -c 1 -o $OUT/${TAG}_full_fwd_bf16 -f \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --extras none > /dev/null 2>$OUT/${TAG}_ncu_err.txt
timeout 200 python scripts/ncu_summary.py $OUT/${TAG}_full_fwd_bf16.ncu-rep 12 > $OUT/${TAG}_ncu_fwd_bf16.txt 2>&1
timeout 100 ncu -i $OUT/${TAG}_full_fwd_bf16.ncu-rep --page details --csv 2>/dev/null | grep -i "pipe\|Executed Ipc\|Issue Slots\|Duration\|DRAM Throughput\|Registers\|Theoretical Occ\|Memory Throughput" | cut -c1-220 >> $OUT/${TAG}_ncu_fwd_bf16.txt
head -4 $OUT/${TAG}_ncu_fwd_bf16.txt | cut -c1-160
rm -f $OUT/*.ncu-rep
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","ms_per_step","gpu_launches") if k in d}, d.get("e2e"), d.get("clocks"), d.get("roofline"))
for k,v in (d.get("extras") or {}).items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","scaling","error")}, (v.get("config") or {}).get("kernel"))
for f in ("reference_arm","flux_causal"):
    try:
        r=json.loads(open("$OUT/${TAG}_bench_%s.json" % f).read().strip().splitlines()[-1]); print(f, round(r["value"],3), r["unit"], r.get("ms_per_step"))
    except Exception as e: print(f, "failed", e)
try:
    d = json.loads(open("$OUT/${TAG}_bench_mask.json").read().strip().splitlines()[-1])
    print({k: (round(v["ms"], 4), round(v.get("bwd_ms", 0), 4), v.get("kernel")) for k, v in d.items() if isinstance(v, dict)})
except Exception as e: print("mask failed", e)
PY
tail -3 $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_ncu_err.txt
