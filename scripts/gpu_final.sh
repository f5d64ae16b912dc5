#!/bin/bash
# Round-end evidence run on one B200: parity tests, bench lines, mask / quantised benches, ncu launch lists and full captures.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_final.sh r01f'
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_gpu_tests.log
timeout 300 python bench.py > $OUT/${TAG}_bench_flux_fwd.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --mode fwdbwd --no-cpu-baseline > $OUT/${TAG}_bench_flux_fwdbwd.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload long_window --mode fwdbwd --steps 10 --no-cpu-baseline > $OUT/${TAG}_bench_c4_fwdbwd.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload flux_causal --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_flux_causal.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload ring128k --steps 4 --warmup 2 > $OUT/${TAG}_bench_causal128k_1gpu.json 2>> $OUT/${TAG}_bench.err
timeout 300 python scripts/bench_quant.py 10 > $OUT/${TAG}_bench_quant.json 2>> $OUT/${TAG}_bench.err
timeout 300 python scripts/bench_mask.py 10 > $OUT/${TAG}_bench_mask.json 2>> $OUT/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fwd.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fwdbwd.csv \
    python bench.py --mode fwdbwd --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_quant.csv \
    python scripts/bench_quant.py 2 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_mask.csv \
    python scripts/bench_mask.py 2 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel|bwd_dkv_tc|bwd_dq_tc' --launch-skip 6 -c 3 \
    -o $OUT/${TAG}_full_tc -f python bench.py --mode fwdbwd --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel' --launch-skip 14 -c 1 \
    -o $OUT/${TAG}_full_fwd_int8 -f python scripts/bench_quant.py 2 > /dev/null 2>&1
ls -la $OUT | tail -30
tail -3 $OUT/${TAG}_bench.err
for f in flux_fwd reference_arm flux_fwdbwd c4_fwdbwd flux_causal causal128k_1gpu; do cut -c1-420 $OUT/${TAG}_bench_$f.json; echo; done
cut -c1-1500 $OUT/${TAG}_bench_quant.json; echo; cat $OUT/${TAG}_bench_mask.json
