#!/bin/bash
# One GPU box visit: parity tests, bench lines, ncu launch lists and --set full captures.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/${TAG}_gpu_tests.log
timeout 300 python bench.py > $OUT/${TAG}_bench_flux_fwd.json 2> $OUT/${TAG}_bench_flux_fwd.err
timeout 300 python bench.py --mode fwdbwd --no-cpu-baseline > $OUT/${TAG}_bench_flux_fwdbwd.json 2>> $OUT/${TAG}_bench_flux_fwd.err
timeout 300 python bench.py --workload long_window --mode fwdbwd --steps 10 --no-cpu-baseline > $OUT/${TAG}_bench_c4_fwdbwd.json 2>> $OUT/${TAG}_bench_flux_fwd.err
timeout 300 python bench.py --workload flux_causal --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_flux_causal.json 2>> $OUT/${TAG}_bench_flux_fwd.err
timeout 300 python scripts/bench_quant.py 10 > $OUT/${TAG}_bench_quant.json 2> $OUT/${TAG}_bench_quant.err
# launch lists (per-launch times are cold-cache + serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fwd.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fwdbwd.csv \
    python bench.py --mode fwdbwd --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_quant.csv \
    python scripts/bench_quant.py 2 > /dev/null 2>&1
# full captures of the dominant kernels (one launch each)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel|bwd_dkv_tc|bwd_dq_tc' --launch-skip 6 -c 3 \
    -o $OUT/${TAG}_full_tc -f python bench.py --mode fwdbwd --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_tcq' --launch-skip 4 -c 1 \
    -o $OUT/${TAG}_full_tcq -f python scripts/bench_quant.py 2 > /dev/null 2>&1
ls -la $OUT
cat $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_bench_flux_fwd.json $OUT/${TAG}_bench_flux_fwdbwd.json $OUT/${TAG}_bench_c4_fwdbwd.json $OUT/${TAG}_bench_quant.json
