#!/bin/bash
# Per-launch cycles vs duration (=> SM clock under load) and pipe activity for a few forward variants.
TAG=${1:-clk}
OUT=gpurun_out
mkdir -p $OUT
M=gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
for P in 0 2 4; do
  MFA_FWD_POLY=$P timeout 300 ncu --metrics $M --clock-control none -k regex:fwd_tc_kernel --launch-skip 8 -c 4 --csv --log-file $OUT/${TAG}_poly$P.csv \
    python bench.py --steps 10 --warmup 2 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
MFA_FWD_POLY=0 timeout 300 ncu --metrics $M --clock-control none -k regex:fwd_tc_kernel -c 40 --csv --log-file $OUT/${TAG}_quant.csv python scripts/bench_quant.py 2 > /dev/null 2>&1
python - <<'PY'
import csv,glob,collections
for f in sorted(glob.glob('gpurun_out/*clk*_*.csv')):
    rows=list(csv.reader(open(f)))
    h=[i for i,r in enumerate(rows) if r and r[0]=='ID']
    if not h: print(f,'no data'); continue
    hd=rows[h[0]]; ki=hd.index('Kernel Name'); mi=hd.index('Metric Name'); vi=hd.index('Metric Value'); ii=hd.index('ID')
    d=collections.OrderedDict()
    for r in rows[h[0]+1:]:
        if len(r)>vi: d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=float(r[vi].replace(',',''))
    print(f)
    for (i,k),m in d.items():
        t=m.get('gpu__time_duration.sum',0); c=m.get('sm__cycles_elapsed.avg',0)
        print(f"  {i:>3} {k:60s} {t/1000:8.1f}us cyc={c:9.0f} clk={c/t if t else 0:5.3f}GHz tensor={m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',0):5.1f}% issue={m.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}% inst={m.get('sm__inst_executed.sum',0)/1e6:7.1f}M xu={m.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',0):5.1f}%")
PY
