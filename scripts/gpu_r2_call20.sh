#!/bin/bash
# round 2: ncu --set full captures of the forward kernel in its bf16, int8 + e4m3 and fp32-split modes (pipe breakdown)
OUT=gpurun_out; mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 4 -c 1 \
    -o $OUT/r02u_full_fwd_bf16 -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --extras none > /dev/null 2>$OUT/r02u_err.txt
# bench_quant order per step: bf16, int8_tensor, int8_block64, int4_block64 (3 warm-up + n timed each): skip the 5 bf16 launches
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 12 -c 1 \
    -o $OUT/r02u_full_fwd_int8f8 -f python scripts/bench_quant.py 2 > /dev/null 2>>$OUT/r02u_err.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 17 -c 1 \
    -o $OUT/r02u_full_fwd_int4f8 -f python scripts/bench_quant.py 2 > /dev/null 2>>$OUT/r02u_err.txt
MFA_FP32_SLICE_KEYS=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 1 -c 1 \
    -o $OUT/r02u_full_fwd_fp32split -f python scripts/bench_fp32.py 2 > /dev/null 2>>$OUT/r02u_err.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $OUT/r02u_launches_fp32.csv python scripts/bench_fp32.py 2 > /dev/null 2>&1
grep -v "^==" $OUT/r02u_launches_fp32.csv | awk -F'","' 'NR>1{print $5, $NF}' | head -40 | cut -c1-160
for n in bf16 int8f8 int4f8 fp32split; do
  timeout 200 python scripts/ncu_summary.py $OUT/r02u_full_fwd_$n.ncu-rep 12 > $OUT/r02u_ncu_fwd_$n.txt 2>&1
  timeout 100 ncu -i $OUT/r02u_full_fwd_$n.ncu-rep --page details --csv 2>/dev/null | grep -i "pipe\|Executed Ipc\|Issue Slots\|Duration\|DRAM Throughput\|Registers\|Theoretical Occ" | cut -c1-220 >> $OUT/r02u_ncu_fwd_$n.txt
done
rm -f $OUT/*.ncu-rep
head -20 $OUT/r02u_ncu_fwd_int8f8.txt | cut -c1-220; tail -3 $OUT/r02u_err.txt
