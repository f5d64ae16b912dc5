#!/bin/bash
# staged-mask forward, dev variant: parity subset (D = 128 bf16 only exists in the variant) + forward-only mask bench
TAG=${1:-r02bc}; V=${2:-devone}
OUT=gpurun_out
mkdir -p $OUT
( MFA_LIBRARY=lib_variants/$V/libMFAFFI.so timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "staged and (shape0 or in_place) or mask_with_causal or mask_matches_exact or tile_skipping" 2>&1 | tail -8 ) > $OUT/${TAG}_tests_$V.log
cat $OUT/${TAG}_tests_$V.log
bash scripts/gpu_r2_call47.sh $TAG $V
