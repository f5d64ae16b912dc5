#!/bin/bash
# Sweeps the exp2-offload fraction of the tcgen05 forward (MFA_FWD_POLY) on the FLUX shape.
for P in 0 1 2 3 4; do
  echo -n "POLY=$P "
  MFA_FWD_POLY=$P python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['achieved'],1), d['clocks'])"
done
