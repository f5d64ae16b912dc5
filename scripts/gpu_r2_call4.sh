#!/bin/bash
export MFA_WATCHDOG=1
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python scripts/ring_emulate.py 131072 32 8 > $OUT/r02e_ring_emulate.txt 2>$OUT/r02e_ring_err.txt; cat $OUT/r02e_ring_emulate.txt; grep "mfa\]" $OUT/r02e_ring_err.txt | head -120
