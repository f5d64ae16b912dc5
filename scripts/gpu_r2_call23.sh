#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python scripts/fwd_trace.py int8 3 2 > $OUT/r02x_trace_int8_bias.txt 2>$OUT/r02x_err.txt
timeout 60 python scripts/fwd_trace_events.py $OUT/fwd_trace_int8_3_q2.txt 10 2 >> $OUT/r02x_trace_int8_bias.txt 2>>$OUT/r02x_err.txt
head -70 $OUT/r02x_trace_int8_bias.txt; tail -3 $OUT/r02x_err.txt
