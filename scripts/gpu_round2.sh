#!/bin/bash
# tests + bench lines only (no ncu)
TAG=${1:-r}
OUT=gpurun_out
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/${TAG}_gpu_tests.log
cat $OUT/${TAG}_gpu_tests.log
timeout 300 python bench.py > $OUT/${TAG}_bench_flux_fwd.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --mode fwdbwd --no-cpu-baseline > $OUT/${TAG}_bench_flux_fwdbwd.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload long_window --mode fwdbwd --steps 10 --no-cpu-baseline > $OUT/${TAG}_bench_c4_fwdbwd.json 2>> $OUT/${TAG}_bench.err
timeout 300 python scripts/bench_quant.py 10 > $OUT/${TAG}_bench_quant.json 2>> $OUT/${TAG}_bench.err
for MB in 2 4 8 16; do echo -n "chunk_mb=$MB "; MFA_PIPELINE_CHUNK_MB=$MB timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e'])"; done > $OUT/${TAG}_e2e_sweep.txt 2>&1
cat $OUT/${TAG}_bench.err | tail -5
cat $OUT/${TAG}_bench_flux_fwd.json $OUT/${TAG}_bench_flux_fwdbwd.json $OUT/${TAG}_bench_c4_fwdbwd.json $OUT/${TAG}_bench_quant.json | cut -c1-1500
cat $OUT/${TAG}_e2e_sweep.txt
