#!/bin/bash
# Evidence run after the staged-mask / dterm work (release build): whole GPU suite, smoke, default bench with every extra, reference arm,
# mask bench, launch list, ncu --set full summaries of the headline forward, the staged-mask forward / backward kernels and the dterm
# kernel, and an A/B of backward ring depths.   gpurun --timeout 1700 -- 'bash scripts/gpu_r2_call51.sh r02bf'
TAG=${1:-r02bf}
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
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras fwdbwd_flux,mask_bf16_dense,int8_block,fp32_flux,d256_fwd > /dev/null 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:'fwd_tc_kernel' --launch-skip 4 -c 1 -o $OUT/${TAG}_full_fwd_bf16 -f \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --extras none > /dev/null 2>$OUT/${TAG}_ncu_err.txt
# bench_mask.py with 2 timed steps: 5 forward launches per case, 7th case = additive bf16 [1,H,S,S]
MFA_BENCH_MASK_FWD_ONLY=1 timeout 300 $NCU -k regex:'fwd_tc_kernel' --launch-skip 33 -c 1 -o $OUT/${TAG}_full_fwd_mask_bf16 -f \
    python scripts/bench_mask.py 2 > /dev/null 2>>$OUT/${TAG}_ncu_err.txt
# backward part of bench_mask.py with 2 timed steps: 4 backward calls per case
timeout 300 $NCU -k regex:'bwd_dkv_tc_kernel' --launch-skip 26 -c 1 -o $OUT/${TAG}_full_dkv_mask_bf16 -f \
    python scripts/bench_mask.py 2 > /dev/null 2>>$OUT/${TAG}_ncu_err.txt
timeout 300 $NCU -k regex:'bwd_dq_tc_kernel' --launch-skip 26 -c 1 -o $OUT/${TAG}_full_dq_mask_bf16 -f \
    python scripts/bench_mask.py 2 > /dev/null 2>>$OUT/${TAG}_ncu_err.txt
timeout 300 $NCU -k regex:'dterm_vec_kernel' --launch-skip 2 -c 1 -o $OUT/${TAG}_full_dterm -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras fwdbwd_flux > /dev/null 2>>$OUT/${TAG}_ncu_err.txt
for n in fwd_bf16 fwd_mask_bf16 dkv_mask_bf16 dq_mask_bf16 dterm; do
  [ -f $OUT/${TAG}_full_$n.ncu-rep ] || { echo "no capture $n"; continue; }
  timeout 200 python scripts/ncu_summary.py $OUT/${TAG}_full_$n.ncu-rep 12 > $OUT/${TAG}_ncu_$n.txt 2>&1
  timeout 100 ncu -i $OUT/${TAG}_full_$n.ncu-rep --page details --csv 2>/dev/null | grep -i "pipe\|Executed Ipc\|Issue Slots\|Duration\|DRAM Throughput\|Registers\|Theoretical Occ\|Memory Throughput" | cut -c1-220 >> $OUT/${TAG}_ncu_$n.txt
  head -8 $OUT/${TAG}_ncu_$n.txt | cut -c1-200
done
rm -f $OUT/*.ncu-rep
for V in main "$@"; do
  [ "$V" = "$TAG" ] && continue
  L=universal-metal-flash-attention_b200/lib/libMFAFFI.so; [ "$V" != main ] && L=lib_variants/$V/libMFAFFI.so
  MFA_LIBRARY=$L timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-e2e --extras fwdbwd_flux > $OUT/${TAG}_bwdring_$V.json 2>> $OUT/${TAG}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bwdring_$V.json").read().strip().splitlines()[-1])
    e = d["extras"]["fwdbwd_flux"]; print("$V", "fwd", round(d["value"], 1), "fwdbwd", round(e["value"], 1), round(e["ms_per_step"], 4))
except Exception as ex: print("$V", "failed", ex)
PY
done
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
