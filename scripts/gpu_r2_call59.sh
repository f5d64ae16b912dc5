#!/bin/bash
# source-level ncu capture of the block-64 quantiser (two-trip, 256-thread CTAs) on a FLUX tensor
TAG=${1:-r02bp}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'quant_span_kernel' --launch-skip 3 -c 1 -o $OUT/${TAG}_full_quant_span -f \
    python scripts/bench_helpers.py 2 > /dev/null 2>$OUT/${TAG}_err.txt
timeout 200 python scripts/ncu_summary.py $OUT/${TAG}_full_quant_span.ncu-rep 30 > $OUT/${TAG}_ncu_quant_span.txt 2>&1
timeout 100 ncu -i $OUT/${TAG}_full_quant_span.ncu-rep --page details --csv 2>/dev/null | grep -i "pipe\|Executed Ipc\|Issue Slots\|Duration\|DRAM Throughput\|Registers\|Theoretical Occ\|Achieved Occ\|Memory Throughput\|L2 Hit\|L1/TEX Hit\|Stall\|Warp Cycles Per Issued\|No Eligible\|Eligible Warps" | cut -c1-200 >> $OUT/${TAG}_ncu_quant_span.txt
rm -f $OUT/*.ncu-rep
head -60 $OUT/${TAG}_ncu_quant_span.txt | cut -c1-190
tail -30 $OUT/${TAG}_ncu_quant_span.txt | cut -c1-190
tail -2 $OUT/${TAG}_err.txt
