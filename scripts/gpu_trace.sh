#!/bin/bash
TAG=${1:-tr}
mkdir -p gpurun_out
( for m in bf16 int8; do for P in 0 2; do timeout 120 python scripts/fwd_trace.py $m $P 2>&1 | tail -8; done; done ) > gpurun_out/${TAG}_trace.txt 2>&1
cat gpurun_out/${TAG}_trace.txt
cp /tmp/fwd_trace_*.txt gpurun_out/ 2>/dev/null
